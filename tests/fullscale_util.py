"""Helpers of the full-size (25.7M x 768) GPU parity tests: an fp64 arbiter that regenerates the synthetic
corpus slab by slab on the device (the CPU oracle cannot hold 79 GB), and the query subset that touches every
128-query tile of a 2514-query batch."""
import numpy as np

N_ROWS = 25_700_592
N_QUERIES = 2514
DIM = 768


def queries_over_all_tiles(n_queries=N_QUERIES, per_tile=13):
    """>= 256 query indices spread over every 128-query tile (first, last and evenly spaced ones of each tile)."""
    sel = []
    for t0 in range(0, n_queries, 128):
        t1 = min(n_queries, t0 + 128)
        sel.extend(sorted(set(np.linspace(t0, t1 - 1, per_tile).astype(int).tolist())))
    return np.asarray(sorted(set(sel)), dtype=np.int64)


def arbiter_topk(q_sel, n_synth_rows, k, planted=None, slab=1_000_000, seed=42, device=0, extra=16):
    """Exact top-k of `q_sel` (CUDA fp32 [m, 768]) over synthetic rows [0, n_synth_rows) followed by the optional
    `planted` rows (CUDA fp32), scored in fp64 with torch and ordered by (score desc, id asc).
    Independent of the engine's search kernels: only the row generator is shared.
    Returns (ref_D [m, k], ref_I [m, k], scores_of): `scores_of(qi, ids)` gives the fp64 score of ids ranked up to
    k + extra - what the comparator needs to accept a substitution at the rank-k boundary that is within tolerance."""
    import torch
    from haconvdr_b200.index import synth_rows_device
    m = q_sel.shape[0]
    kk = k + extra
    best_s = torch.full((m, kk), -float("inf"), dtype=torch.float64, device=q_sel.device)
    best_i = torch.full((m, kk), -1, dtype=torch.int64, device=q_sel.device)
    q64 = q_sel.double()
    slabs = [(r0, min(slab, n_synth_rows - r0), False) for r0 in range(0, n_synth_rows, slab)]
    if planted is not None and planted.shape[0]:
        slabs.append((n_synth_rows, planted.shape[0], True))
    for r0, n, is_planted in slabs:
        x = planted if is_planted else synth_rows_device(n, DIM, seed=seed, row0=r0, device=device)
        s = q64 @ x.double().T
        s, i = torch.topk(s, min(kk, n), dim=1)
        cat_s = torch.cat([best_s, s], 1)
        cat_i = torch.cat([best_i, i + r0], 1)
        top = torch.topk(cat_s, kk, dim=1)
        best_s, best_i = top.values, torch.gather(cat_i, 1, top.indices)
        del x, s
    ext_D, ext_I = best_s.cpu().numpy(), best_i.cpu().numpy()
    order = np.lexsort((ext_I, -ext_D), axis=1)
    ext_D, ext_I = np.take_along_axis(ext_D, order, 1), np.take_along_axis(ext_I, order, 1)
    lookup = [dict(zip(ext_I[r].tolist(), ext_D[r].tolist())) for r in range(m)]

    def scores_of(qi, ids):
        return np.asarray([lookup[qi].get(int(i), -np.inf) for i in ids], dtype=np.float64)

    return ext_D[:, :k].copy(), ext_I[:, :k].copy(), scores_of
