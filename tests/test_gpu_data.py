"""GPU parity of the device-resident window dataset (csrc/windows.cu, gdn_b200/datasets) against the
reference's own TimeDataset vectors and the oracle.  Bit-exact (float32 copies)."""
import numpy as np
import pytest
import torch

from golden_util import load
from oracle import data_oracle as do

pytestmark = pytest.mark.gpu
CASES = [("train", 5, 3), ("test", 5, 3), ("train", 1, 1), ("test", 16, 7)]


def _make(mode, W, S, raw):
    from gdn_b200.datasets import TimeDataset
    return TimeDataset(raw, torch.zeros(2, 4, dtype=torch.long), mode=mode, config={"slide_win": W, "slide_stride": S})


@pytest.mark.parametrize("mode,W,S", CASES)
def test_windows_equal_reference_vectors(mode, W, S):
    rec = load("timedataset_small")
    ds = _make(mode, W, S, rec["raw"].tolist())
    tag = f"{mode}_w{W}_s{S}"
    assert len(ds) == rec[tag + "_x"].shape[0]
    x, y, lab = ds.batch(torch.arange(len(ds)))
    assert np.array_equal(x.cpu().numpy(), rec[tag + "_x"].astype(np.float32))     # train.py:66 casts to float
    assert np.array_equal(y.cpu().numpy(), rec[tag + "_y"].astype(np.float32))
    assert np.array_equal(lab.cpu().numpy(), rec[tag + "_labels"])
    assert np.array_equal(ds.labels.numpy(), rec[tag + "_labels"])
    item = ds[2]                                                                   # reference item protocol
    assert item[0].dtype == torch.float64 and np.array_equal(item[0].numpy(), rec[tag + "_item2_x"])
    assert np.array_equal(item[1].numpy(), rec[tag + "_item2_y"]) and item[3].dtype == torch.long
    assert np.array_equal(ds[-1][0].numpy(), rec[tag + "_x"][-1])


def test_loader_batches_subsets_and_shuffle():
    rng = np.random.default_rng(3)
    N, T, W = 70, 400, 12
    raw = np.concatenate([rng.random((N, T)), np.zeros((1, T))], 0).tolist()
    ds = _make("train", W, 5, raw)
    xo, yo, _ = do.process(raw, W, 5, "train")
    seen = []
    for x, y, lab, ei in ds.loader(32):                                            # ragged last batch
        assert x.is_cuda and x.dtype == torch.float32 and ei is ds.edge_index
        seen.append((x.cpu().numpy(), y.cpu().numpy()))
    assert len(seen) == len(ds.loader(32)) == (len(ds) + 31) // 32
    assert np.array_equal(np.concatenate([s[0] for s in seen]), xo.astype(np.float32))
    assert np.array_equal(np.concatenate([s[1] for s in seen]), yo.astype(np.float32))
    sub = [5, 1, 17, 3]                                                            # Subset semantics (main.py:71-79)
    (x, y, lab, _), = list(ds.loader(8, indices=sub))
    assert np.array_equal(x.cpu().numpy(), xo[sub].astype(np.float32))
    g = torch.Generator(device="cuda").manual_seed(0)
    got = torch.cat([b[1] for b in ds.loader(16, shuffle=True, generator=g)]).cpu().numpy()
    assert got.shape == yo.shape and not np.array_equal(got, yo.astype(np.float32))
    assert np.array_equal(np.sort(got.sum(1)), np.sort(yo.astype(np.float32).sum(1)))   # a permutation of the windows
    assert len(list(ds.loader(16, drop_last=True))) == len(ds) // 16


def test_edges_and_errors():
    from gdn_b200 import _lib
    from gdn_b200._lib import ptr
    raw = [[float(t) for t in range(9)], [0.0] * 9]
    ds = _make("test", 8, 1, raw)                                                  # a single window: e = 8 = T - 1
    assert len(ds) == 1
    x, y, _ = ds.batch([0])
    assert x.flatten().tolist() == [float(t) for t in range(8)] and y.item() == 8.0
    x, y, lab = ds.batch([])
    assert x.shape == (0, 1, 8) and y.shape == (0, 1)
    with pytest.raises(IndexError):
        ds.batch([1])
    with pytest.raises(IndexError):
        ds.loader(4, indices=[0, 2])
    assert len(_make("test", 9, 1, raw)) == 0                                       # slide_win == T: no window
    # the kernel itself flags a window that leaves the series (C ABI contract)
    lib = _lib.load()
    ends = torch.tensor([8, 3, 9], dtype=torch.int32, device="cuda")               # 3 < W and 9 >= T are invalid
    xb = torch.full((3, 1, 8), -1.0, device="cuda"); yb = torch.full((3, 1), -1.0, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    rc = lib.gdn_window_batch(ptr(ds.series), None, 1, 9, 8, ptr(ends), 3, ptr(xb), ptr(yb), None, ptr(err),
                              torch.cuda.current_stream().cuda_stream)
    # ... and hands zeros, never uninitialised memory, to whoever ignores the flag
    assert rc == 0 and err.item() in (2, 3) and yb.flatten().tolist() == [8.0, 0.0, 0.0]
    assert xb[1:].abs().max().item() == 0.0
    assert lib.gdn_window_batch(ptr(ds.series), None, 1, 9, 9, ptr(ends), 3, ptr(xb), ptr(yb), None, ptr(err), None) < 0


def test_full_size_property_windows_are_views_of_the_series():
    """C5-sized feed: every gathered window equals the strided view of the series (size-independent check)."""
    N, T, W, B = 16384, 512, 16, 64
    series = torch.rand(N, T, device="cuda")
    from gdn_b200.datasets import TimeDataset
    ds = TimeDataset.from_series(series, None, None, mode="test", config={"slide_win": W, "slide_stride": 1})
    assert len(ds) == T - W
    idx = torch.randint(0, T - W, (B,), device="cuda")
    x, y, _ = ds.batch(idx)
    view = series.unfold(1, W + 1, 1)                                              # [N, T-W, W+1]
    want = view[:, idx, :].permute(1, 0, 2)
    assert torch.equal(x, want[..., :W]) and torch.equal(y, want[..., W])


def test_test_loop_matches_reference():
    """gdn_b200.test.test (test.py:20-75) over the device-resident loader, reference weights loaded: the
    reference's avg_loss / predictions within 1e-4, ground truth and labels exact; the result feeds the scorer
    without leaving the device and equals scoring the reference's lists."""
    from gdn_b200.datasets import TimeDataset
    from gdn_b200.evaluate import get_full_err_scores
    from gdn_b200.models.GDN import GDN
    from gdn_b200.test import test
    rec = load("test_loop_small")
    N, T, W, K, D, B = (int(v) for v in rec["dims"])
    ei = torch.zeros(2, 1, dtype=torch.long)
    ds = TimeDataset(rec["raw"].tolist(), ei, mode="test", config={"slide_win": W, "slide_stride": 1})
    model = GDN([ei], N, dim=D, input_dim=W, topk=K)
    model.load_state_dict({k[3:]: torch.from_numpy(v) for k, v in rec.items() if k.startswith("sd.")})
    model = model.cuda()
    avg, res = test(model, ds.loader(B))
    assert abs(avg - float(rec["avg_loss"])) <= 1e-4 * abs(float(rec["avg_loss"]))
    assert len(res) == 3 and res[0].shape == rec["pred"].shape and not res[0].is_cuda
    assert np.abs(res[0].numpy() - rec["pred"]).max() <= 1e-4 * np.abs(rec["pred"]).max()
    assert np.array_equal(res[1].numpy(), rec["gt"]) and np.array_equal(res[2].numpy(), rec["labels"])
    assert np.asarray(res).shape == (3,) + rec["pred"].shape                       # main.py:get_score does np.array(...)
    # a second pass with ragged batches, and the reference's DataLoader protocol (CPU doubles per item)
    avg2, res2 = test(model, ds.loader(7))
    assert torch.equal(res2[1], res[1]) and np.abs(res2[0].numpy() - res[0].numpy()).max() <= 1e-6
    dl = torch.utils.data.DataLoader(ds, batch_size=B, shuffle=False)
    avg3, res3 = test(model, dl)
    assert abs(avg3 - avg) <= 1e-6 * abs(avg) and torch.equal(res3[1], res[1])
    s_dev, n_dev = get_full_err_scores(res, res)
    s_ref, _ = get_full_err_scores([res[0].tolist(), res[1].tolist(), res[2].tolist()], [rec["pred"], rec["gt"], rec["labels"]])
    assert np.array_equal(s_dev, s_ref) and np.array_equal(s_dev, n_dev)
    with pytest.raises(RuntimeError, match="no batch"):
        test(model, [])


def test_reference_call_sequence_end_to_end():
    """The call sequence of main.py / train.py / test.py on top of the mirrors, on a B200: a torch DataLoader over the
    reference-style dataset items (pageable float64, main.py:84-85) feeding train.py:63-77's step verbatim, then
    test.py's loop, evaluate.py's scores and the summary metrics (main.py:150-174).  Checks the plumbing, not
    accuracy: losses fall, every stage returns finite values of the reference's shapes and types."""
    import torch.nn.functional as F
    from torch.utils.data import DataLoader, Dataset
    from gdn_b200 import evaluate as ev
    from gdn_b200.data import Prefetcher
    from gdn_b200.models.GDN import GDN
    from gdn_b200.test import test as gdn_test
    rng = np.random.default_rng(1)
    N, T, W, K, D = 27, 700, 5, 5, 64
    t = np.arange(T)
    series = np.clip(0.5 + 0.3 * np.sin(t[None, :] * rng.uniform(0.02, 0.2, (N, 1)) + rng.uniform(0, 6, (N, 1)))
                     + rng.normal(0, 0.02, (N, T)), 0, 1)
    labels = np.zeros(T)
    labels[500:560] = 1
    series[3, 500:560] += 0.6                                           # an attack on one sensor
    fc = torch.tensor([[j for i in range(N) for j in range(N) if i != j],
                       [i for i in range(N) for j in range(N) if i != j]], dtype=torch.long)   # util/net_struct.py

    class Windows(Dataset):                                             # datasets/TimeDataset.py:33-73 item protocol
        def __init__(self, lo, hi, stride):
            self.ends = list(range(max(lo, W), hi, stride))

        def __len__(self):
            return len(self.ends)

        def __getitem__(self, k):
            e = self.ends[k]
            return (torch.from_numpy(series[:, e - W:e]).double(), torch.from_numpy(series[:, e]).double(),
                    torch.tensor(labels[e]).double(), fc.long())

    train_loader = DataLoader(Windows(0, 450, 1), batch_size=32, shuffle=True)
    test_loader = DataLoader(Windows(450, T, 1), batch_size=32, shuffle=False)
    device = torch.device("cuda")
    torch.manual_seed(5)
    model = GDN([fc], N, dim=D, input_dim=W, out_layer_num=1, out_layer_inter_dim=128, topk=K).to(device)
    optimizer = torch.optim.Adam(model.parameters(), lr=0.001, weight_decay=0)
    epoch_losses = []
    for i_epoch in range(3):
        acu_loss = 0.0
        model.train()
        for x, lab, attack_labels, edge_index in train_loader:          # train.py:63-77
            x, lab, edge_index = [item.float().to(device) for item in [x, lab, edge_index]]
            optimizer.zero_grad()
            out = model(x, edge_index).float().to(device)
            loss = F.mse_loss(out, lab, reduction="mean")
            loss.backward()
            optimizer.step()
            acu_loss += loss.item()
        epoch_losses.append(acu_loss / len(train_loader))
    assert epoch_losses[-1] < epoch_losses[0] and np.isfinite(epoch_losses).all()
    assert tuple(model.learned_graph.shape) == (N, K) and model.learned_graph.dtype == torch.int64
    avg_loss, result = gdn_test(model, Prefetcher(test_loader, device))          # test.py:21-79 (batches prefetched)
    assert np.isfinite(avg_loss) and len(result) == 3 and tuple(result[0].shape) == (len(test_loader.dataset), N)
    scores, normals = ev.get_full_err_scores(result, result)                      # main.py:150-158
    assert scores.shape == (N, len(test_loader.dataset)) and scores.dtype == np.float64 and np.isfinite(scores).all()
    gt_labels = np.asarray(result[2])[:, 0].tolist()
    info = ev.get_best_performance_data(scores, gt_labels, topk=1)                # main.py:164-174
    assert len(info) == 5 and 0.0 <= info[0] <= 1.0 and all(np.isfinite(v) for v in info[:4])
    top_sensor = int(np.argmax(scores[:, 55:100].max(axis=1)))                    # the attacked sensor stands out
    assert top_sensor == 3


def test_sharded_test_slices_cover_the_run():
    """gdn_b200.dp.sharded_test / LoaderSlice (row e-2): the per-rank slices of the evaluation loader -- the
    device-resident WindowLoader (re-indexed) and a plain list of host batches (skipped through) -- concatenate to the
    single-process test() result; with one process sharded_test IS test()."""
    from gdn_b200.dp import LoaderSlice, shard_bounds, sharded_test
    from gdn_b200.models.GDN import GDN
    from gdn_b200.test import test as gdn_test
    rng = np.random.default_rng(5)
    N, T, W = 27, 300, 5
    raw = np.concatenate([rng.random((N, T)), (rng.random((1, T)) > 0.9).astype(np.float64)], 0).tolist()
    ds = _make("test", W, 1, raw)
    torch.manual_seed(1)
    model = GDN([torch.zeros(2, 1, dtype=torch.long)], N, dim=64, input_dim=W, topk=5).cuda().eval()
    loader = ds.loader(32)
    full_loss, full = gdn_test(model, loader)
    for world in (2, 3):
        preds, gts, labs = [], [], []
        for r in range(world):
            lo, hi = shard_bounds(len(loader), r, world)
            _, res = gdn_test(model, LoaderSlice(loader, lo, hi))
            preds.append(res.device_tensors[0]); gts.append(res.device_tensors[1]); labs.append(res.device_tensors[2])
        assert torch.equal(torch.cat(preds), full.device_tensors[0])
        assert torch.equal(torch.cat(gts), full.device_tensors[1])
        assert torch.equal(torch.cat(labs), full.device_tensors[2])
    host = [tuple(t.cpu() if torch.is_tensor(t) else t for t in b) for b in loader]
    _, part = gdn_test(model, LoaderSlice(host, 2, 5))
    assert torch.equal(part.device_tensors[0], full.device_tensors[0][64:160])
    loss1, p1, g1, l1 = sharded_test(model, loader)
    assert abs(loss1 - full_loss) < 1e-9 and torch.equal(p1, full.device_tensors[0]) and torch.equal(l1, full.device_tensors[2][:, 0])
