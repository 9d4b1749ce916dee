"""Top-k comparison protocol for the learned graph ("Tolerance protocol").

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference ranks fp32 cosines that come out of a library SGEMM (models/GDN.py:150-157); any other
exact-arithmetic evaluation order moves each cosine by a few 1e-8, and `torch.topk` breaks exact ties
in an order that is an artefact of libstdc++'s partial_sort / nth_element --
not lowest-index-first.  So:

  * a row is CLEAN when the reference's top-(K+1) cosines are pairwise more than `tau` apart
    (adjacent gaps of the sorted values > tau).  Clean rows must match `torch.topk` bit for bit,
    order included;
  * every other row is TIE-AFFECTED.  Sorted positions whose values chain together with gaps <= tau
    form a tau-cluster; two rankings are canonically equal when they pick the same NUMBER of members
    from every cluster, in cluster order (inside a cluster any order, and at the cluster that
    straddles position K any members).

`compare_topk` returns the counts; tests assert `clean_mismatch == 0 and tied_fail == 0` and print
the rest, bench.py records them.
"""
import torch

TAU = 1e-6


def reference_cosines(V):
    """models/GDN.py:145-152 (same op sequence as oracle.gdn_oracle.learned_graph)."""
    w = V.detach().clone().float()
    gram = torch.matmul(w, w.T)
    nrm = w.norm(dim=-1)
    return gram / torch.matmul(nrm.view(-1, 1), nrm.view(1, -1))


def compare_topk(cos_ref, idx_ours, K, tau=TAU, slack=24):
    """cos_ref [N, N] float32 (the reference's cosine matrix), idx_ours [N, K] int64.

    Returns a dict of counts:
      rows, clean, clean_exact, clean_mismatch, tied, tied_exact, tied_canonical, tied_fail,
      unresolved (boundary cluster longer than the inspected window; counted in tied_fail),
      exact_rows (bit-equal to torch.topk incl. order), exact_ties (adjacent pairs with gap == 0
      inside the top-(K+1)), pairs_within_tau (adjacent pairs with gap <= tau inside the top-(K+1)).
    """
    cos_ref = torch.as_tensor(cos_ref)
    idx_ours = torch.as_tensor(idx_ours).long().cpu()
    N = cos_ref.shape[0]
    K = int(K)
    ext = min(N, K + int(slack))
    idx_ref = torch.topk(cos_ref, K, dim=-1)[1]                     # what the reference publishes
    vals, ids = torch.topk(cos_ref, ext, dim=-1)                    # sorted descending
    gap = vals[:, :-1] - vals[:, 1:]                                # [N, ext-1], >= 0
    close = gap <= tau
    head = close[:, :min(K, ext - 1)]                               # gaps among the first K+1 values
    clean = ~head.any(dim=1)
    exact = (idx_ours == idx_ref).all(dim=1)
    out = {
        "rows": N, "tau": float(tau),
        "clean": int(clean.sum()), "clean_exact": int((clean & exact).sum()),
        "clean_mismatch": int((clean & ~exact).sum()),
        "tied": int((~clean).sum()), "tied_exact": int((~clean & exact).sum()),
        "exact_rows": int(exact.sum()),
        "exact_ties": int((gap[:, :min(K, ext - 1)] == 0).sum()),
        "pairs_within_tau": int(head.sum()),
    }
    rows = torch.nonzero(~clean & ~exact).flatten()
    canonical = 0
    unresolved = 0
    if rows.numel():
        cid = torch.zeros((rows.numel(), ext), dtype=torch.long)
        cid[:, 1:] = torch.cumsum((~close[rows]).long(), dim=1)     # cluster id of every sorted position
        ours = idx_ours[rows]                                        # [R, K]
        eq = ids[rows].unsqueeze(1) == ours.unsqueeze(2)             # [R, K, ext]
        found = eq.any(dim=2)
        pos = eq.float().argmax(dim=2)
        our_cid = cid.gather(1, pos)
        ref_cid = cid[:, :K]
        ordered = (our_cid[:, 1:] >= our_cid[:, :-1]).all(dim=1)
        same_counts = (torch.sort(our_cid, dim=1)[0] == ref_cid).all(dim=1)
        ok = found.all(dim=1) & ordered & same_counts
        if ext < N:
            # the cluster straddling position K must end inside the window, else membership is unknown
            tail_open = close[rows][:, K - 1:].all(dim=1) if K - 1 < ext - 1 else torch.zeros(rows.numel(), dtype=torch.bool)
            unresolved = int((tail_open & ~ok).sum())
        canonical = int(ok.sum())
    out["tied_canonical"] = canonical
    out["tied_fail"] = int(rows.numel()) - canonical
    out["unresolved"] = unresolved
    out["ok"] = out["clean_mismatch"] == 0 and out["tied_fail"] == 0
    return out


def compare_topk_from_embedding(V, idx_ours, K, tau=TAU):
    return compare_topk(reference_cosines(torch.as_tensor(V).cpu()), idx_ours, K, tau=tau)
