"""Import the reference's own hot-path functions with their missing third-party
imports stubbed (TEST INFRASTRUCTURE ONLY; usable only where /root/reference is
mounted, i.e. the build container - never at test/bench run time on the GPU box).

`/root/reference/src/test_HAConvDR_topiocqa.py` imports faiss, pytrec_eval,
IPython, and sibling modules that do not import in this image.
None of them is used by ``search_one_by_one_with_faiss`` (`:74-162`) or by the
mapping loop of ``output_test_res`` (`:232-255`) beyond the ``index`` object the
caller passes in, so they are replaced by empty stand-in modules.  The functions
that run are the reference's own, unmodified.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("HAC_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "test_HAConvDR_topiocqa.py"))


def load_reference_module(name: str = "test_HAConvDR_topiocqa"):
    if not available():
        raise RuntimeError("reference tree not mounted at %s" % REFERENCE_ROOT)
    stubs = {}

    def stub(modname, **attrs):
        m = types.ModuleType(modname)
        for k, v in attrs.items():
            setattr(m, k, v)
        stubs[modname] = m

    dummy = type("Dummy", (), {})
    stub("IPython", embed=lambda *a, **k: None)
    stub("faiss")
    stub("pytrec_eval")
    stub("models", ANCE=dummy, load_model=None)
    stub("utils", check_dir_exist_or_build=None, pstore=None, pload=None, set_seed=None, get_optimizer=None,
         split_and_padding_neighbor=None)
    stub("preprocess_topiocqa", load_collection=None)
    stub("preprocess_qrecc", load_collection=None)
    stub("toml")
    stub("data", padding_seq_to_same_length=None, Retrieval_topiocqa=dummy, Retrieval_topiocqa_old=dummy,
         Retrieval_qrecc=dummy, Retrieval_qrecc_old=dummy, Retrieval_qrecc_new=dummy,
         Retrieval_topiocqa_new=dummy, ConvDataset=dummy, ConvDataset_topiocqa=dummy,
         ConvDataset_topiocqa_rel=dummy, ConvDataset_qrecc=dummy, ConvDataset_qrecc_rel=dummy)
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        path = os.path.join(REFERENCE_ROOT, "src", name + ".py")
        spec = importlib.util.spec_from_file_location("_hac_ref_" + name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    mod.logger.setLevel("WARNING")
    return mod
