"""vast_b200 -- B200-native (sm_100a) cross-modal contrastive + retrieval-scoring hot path of VAST.

Host side is Python/PyTorch (device memory, streams, torch.distributed); all compute on the path runs in
hand-written CUDA behind the C-ABI of include/vast_b200.h (vast_b200/_C/libvast_b200.so, loaded with
ctypes).  There is no CPU or PyTorch fallback: ops raise if the library or a CUDA device is missing."""
from . import ops  # noqa: F401
from .contrastive import forward_ret, gather_negatives, omc_loss_and_negatives  # noqa: F401
from .distributed import (all_gather_ids, all_gather_list, all_gather_with_grad, any_broadcast,  # noqa: F401
                          concat_all_gather, ddp_allgather, exchange_rows)
from .graphed import OmcGraphStep  # noqa: F401
from .features import (batch_get, build_feature, compute_slice_scores, install, l2_normalize, match_head_scores,  # noqa: F401
                       pool_concat, project_normalize)
from .retrieval import (compute_metric_ret, evaluate_ret, rank_of_gt, recall_from_candidates, recall_from_feats,  # noqa: F401
                        refine_candidates, refine_score_matrix, retrieval_topk)

__version__ = "0.1.0"
