from .TimeDataset import TimeDataset, WindowLoader  # noqa: F401
