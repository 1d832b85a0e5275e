"""B200-native exact inner-product search engine - drop-in for the faiss IndexFlatIP
hot path of fengranMark/HAConvDR (src/test_HAConvDR_{topiocqa,qrecc}.py:39-162).

Public surface:
  FlatIPIndex            add / search / reset / ntotal / d  (one GPU shard, C-ABI CUDA library)
  ShardedFlatIPIndex     same surface over torch.distributed ranks (one shard per GPU)
  faiss_compat           stand-ins for the faiss names the reference driver touches
  retrieval              search_one_by_one (reference loop), search_resident, offset2pid, TREC writer
  loader                 block-pickle loader (pinned staging -> HBM)
"""
from ._lib import HAC_MAX_K, HAC_PATH_AUTO, HAC_PATH_GEMV, HAC_PATH_I8, HAC_PATH_MMA  # noqa: F401
from .index import FlatIPIndex  # noqa: F401

__all__ = ["FlatIPIndex", "HAC_PATH_AUTO", "HAC_PATH_GEMV", "HAC_PATH_MMA", "HAC_PATH_I8", "HAC_MAX_K"]
