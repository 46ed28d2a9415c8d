"""Host-side logic of the streaming evaluation (vast_b200/retrieval.py) on CPU tensors against the oracle:
`recall_from_candidates` ranks the ground truth in the matrix `refine_score_matrix` (evaluation_mm.py:253-319) would
have built -- zeros except the ITM scores at the candidate positions -- without building it; the oracle's
`compute_metric_ret` (evaluation_mm.py:326-380 restated, pinned to the reference by tests/golden) evaluates the
materialised matrix.  No CUDA kernels are involved (pure index arithmetic), so these run everywhere."""
import numpy as np
import pytest
import torch

from oracle import spec
from vast_b200 import retrieval


def _random_lists(rng, rows, cols, k, zero_frac=0.0, tie_frac=0.0):
    """idx [rows, k] distinct columns per row (-1 padding when k > cols), itm scores in (0, 1)."""
    kk = min(k, cols)
    idx = np.full((rows, k), -1, dtype=np.int32)
    itm = np.zeros((rows, k), dtype=np.float32)
    for r in range(rows):
        idx[r, :kk] = rng.permutation(cols)[:kk]
        s = rng.random(kk).astype(np.float32) * 0.98 + 0.01
        if tie_frac:
            s = np.where(rng.random(kk) < tie_frac, np.float32(0.5), s)      # exact ties between candidates
        if zero_frac:
            s = np.where(rng.random(kk) < zero_frac, np.float32(0.0), s)     # scores that underflowed to 0
        itm[r, :kk] = s
    return idx, itm


def _dense(idx, itm, rows, cols, transpose=False):
    """[rows, cols] matrix; transpose: the lists are per COLUMN and hold row indices."""
    d = np.zeros((rows, cols), dtype=np.float32)
    for r in range(idx.shape[0]):
        for j in range(idx.shape[1]):
            if idx[r, j] >= 0:
                if transpose:
                    d[idx[r, j], r] = itm[r, j]
                else:
                    d[r, idx[r, j]] = itm[r, j]
    return d


@pytest.mark.parametrize("nt,nv,k,per", [(60, 12, 16, 5), (40, 40, 3, 1), (35, 7, 1, 5), (50, 25, 9, 2), (30, 30, 50, 1)])
@pytest.mark.parametrize("zero_frac,tie_frac", [(0.0, 0.0), (0.3, 0.3)])
def test_forward_metrics_from_lists_equal_dense_oracle(nt, nv, k, per, zero_frac, tie_frac):
    rng = np.random.default_rng(nt * 131 + k)
    idx, itm = _random_lists(rng, nt, nv, k, zero_frac, tie_frac)
    ids = [f"v{i}" for i in range(nv)]
    ids_txt = [f"v{(i // per) % nv}" for i in range(nt)]
    got = retrieval.recall_from_candidates(torch.from_numpy(idx), torch.from_numpy(itm), ids, ids_txt, "forward")
    assert got == spec.compute_metric_ret(_dense(idx, itm, nt, nv), ids, ids_txt, "forward")


@pytest.mark.parametrize("nt,nv,k,per", [(60, 12, 16, 5), (40, 40, 3, 1), (35, 7, 1, 5), (50, 25, 9, 2)])
@pytest.mark.parametrize("zero_frac,tie_frac", [(0.0, 0.0), (0.3, 0.3)])
def test_backward_metrics_from_lists_equal_dense_oracle(nt, nv, k, per, zero_frac, tie_frac):
    """backward: lists are per VIDEO (k best texts); the metric is the min rank over a video's captions."""
    rng = np.random.default_rng(nv * 17 + k)
    idx, itm = _random_lists(rng, nv, nt, k, zero_frac, tie_frac)
    ids = [f"v{i}" for i in range(nv)]
    ids_txt = [f"v{(i // per) % nv}" for i in range(nt)]
    got = retrieval.recall_from_candidates(torch.from_numpy(idx), torch.from_numpy(itm), ids, ids_txt, "backward")
    assert got == spec.compute_metric_ret(_dense(idx, itm, nt, nv, transpose=True), ids, ids_txt, "backward")


def test_rank_in_sparse_row_matches_rank_of_gt():
    rng = np.random.default_rng(5)
    rows, cols, k = 200, 33, 6
    idx, itm = _random_lists(rng, rows, cols, k, zero_frac=0.2, tie_frac=0.2)
    gt = rng.integers(0, cols, rows)
    got = retrieval._rank_in_sparse_row(torch.from_numpy(idx), torch.from_numpy(itm), torch.from_numpy(gt)).numpy()
    assert np.array_equal(got, spec.rank_of_gt(_dense(idx, itm, rows, cols), gt))


def test_duplicate_video_ids_use_first_occurrence():
    """`ids.index(...)` semantics (evaluation_mm.py:337): with duplicate video ids the FIRST column is the ground truth."""
    ids = ["a", "b", "a", "c"]
    ids_txt = ["a", "c", "b"]
    idx = torch.tensor([[2, 0], [3, 1], [1, 2]], dtype=torch.int32)
    itm = torch.tensor([[0.9, 0.8], [0.7, 0.2], [0.6, 0.5]])
    got = retrieval.recall_from_candidates(idx, itm, ids, ids_txt, "forward")
    assert got == spec.compute_metric_ret(_dense(idx.numpy(), itm.numpy(), 3, 4), ids, ids_txt, "forward")
    assert got["forward_r1"] == round(2 / 3 * 100, 1)
