"""CPU: the C-ABI library loads and exports every symbol include/hac_index.h declares; host-side
logic (pickle header walker, reference-loop mirror, TREC writer) without touching a GPU."""
import os
import pickle
import re
import tempfile

import numpy as np
import pytest

from haconvdr_b200 import _lib, loader, retrieval
from oracle.flat_ip import FlatIP, offsets_to_ranked_pids, search_one_by_one, trec_lines
from helpers import GOLDEN_MERGE_CASES, load_golden, write_blocks

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "hac_index.h")).read()
    declared = set(re.findall(r"\b(hac_[a-z_]+)\s*\(", hdr))
    assert declared == set(_lib.SIGNATURES)
    L = _lib.lib()                                  # binds all of them (AttributeError otherwise)
    assert L.hac_abi_version() == _lib.HAC_ABI_VERSION
    assert isinstance(_lib.last_error(), str)


def test_shard_group_rejects_bad_arguments_without_a_gpu():
    """hac_shards_* argument checks come before any CUDA call (error codes and messages as every other entry point)."""
    import ctypes
    L = _lib.lib()
    grp = ctypes.c_void_p()
    assert L.hac_shards_create(None, 0, ctypes.byref(grp)) == -1 and "shards" in _lib.last_error()
    handles = (ctypes.c_void_p * 2)(None, None)
    assert L.hac_shards_create(handles, 2, ctypes.byref(grp)) == -1 and not grp.value
    assert L.hac_shards_create(handles, 2, None) == -1
    assert L.hac_shards_search(None, 1, None, 10, None, None) == -1
    assert L.hac_shards_set_exchange(None, 1) == -1
    assert L.hac_shards_peer_access(None) == 0
    assert L.hac_shards_destroy(None) == 0


def test_library_is_sm100a_with_tcgen05_and_bulk_copy():
    import subprocess
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    if not sass:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in sass
    assert "UTCHMMA" in sass and "LDTM" in sass and "UBLKCP" in sass
    # the default scan: int8 tensor-core MMA issued by CTA pairs, 3-input max in the epilogue
    assert "UTCIMMA.2CTA" in sass and "VIMNMX3" in sass and "FMNMX3" in sass


@pytest.mark.parametrize("proto", [3, 4])
@pytest.mark.parametrize("shape", [(1000, 768), (3, 64), (1, 768)])
def test_pickle_header_walker(proto, shape):
    a = np.random.default_rng(1).standard_normal(shape).astype(np.float32)
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "passage_emb_block_0.pb")
        with open(p, "wb") as h:
            pickle.dump(a, h, protocol=proto)
        hdr = loader.parse_ndarray_pickle_header(p)
        assert hdr.shape == shape and hdr.dtype == np.dtype("<f4")
        raw = open(p, "rb").read()[hdr.payload_offset: hdr.payload_offset + hdr.payload_bytes]
        assert np.array_equal(np.frombuffer(raw, np.float32).reshape(shape), a)
        # numpy 1.x module path ("numpy.core.multiarray"), as written by the reference-era stack (numpy 1.22): the
        # same stream with the module string one byte shorter (length byte / FRAME length patched, so it still
        # unpickles) must give the same header
        raw = open(p, "rb").read()
        if proto >= 4:
            old_s, new_s = b"\x8c\x16numpy._core.multiarray", b"\x8c\x15numpy.core.multiarray"
            assert old_s in raw
            at = raw.index(old_s)
            raw1 = raw.replace(old_s, new_s, 1)
            assert raw1[2:3] == b"\x95"                          # FRAME: its length shrinks by the removed byte
            flen = int.from_bytes(raw1[3:11], "little")
            if at < 11 + flen:
                raw1 = raw1[:3] + (flen - 1).to_bytes(8, "little") + raw1[11:]
        else:
            raw1 = raw.replace(b"cnumpy._core.multiarray\n", b"cnumpy.core.multiarray\n", 1)
        assert raw1 != raw and b"numpy.core.multiarray" in raw1
        p1 = os.path.join(d, "numpy1_era.pb")
        with open(p1, "wb") as h:
            h.write(raw1)
        assert np.array_equal(pickle.load(open(p1, "rb")), a)      # a valid pickle (numpy keeps the old module path alive)
        hdr1 = loader.parse_ndarray_pickle_header(p1)
        assert hdr1.shape == shape and hdr1.dtype == hdr.dtype and hdr1.payload_bytes == hdr.payload_bytes
        assert hdr1.payload_offset == hdr.payload_offset - 1
        got = open(p1, "rb").read()[hdr1.payload_offset: hdr1.payload_offset + hdr1.payload_bytes]
        assert np.array_equal(np.frombuffer(got, np.float32).reshape(shape), a)
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "not_an_array.pb")
        with open(p, "wb") as h:
            pickle.dump({"a": 1}, h, protocol=4)
        with pytest.raises(ValueError):
            loader.parse_ndarray_pickle_header(p)


def test_pickle_header_walker_binbytes8_payload_beyond_4gib():
    """Every real block is 7.68 GB (`/root/reference/gen_doc_embeddings.py:87-88`, 2.5M x 768 fp32): its payload is
    a BINBYTES8 item.  A sparse 4.3 GB file with the opcode stream of such a pickle (the header of a small ndarray
    pickle with the shape and the payload item rewritten) must parse, and rows must be found where they are."""
    n, d = 1_400_001, 768
    small = np.zeros((2, d), np.float32)
    raw = pickle.dumps(small, protocol=4)
    shape_ops = b"K\x02M\x00\x03\x86"                           # (2, 768): BININT1 2, BININT2 768, TUPLE2
    assert raw.count(shape_ops) == 1
    big_shape = b"J" + int(n).to_bytes(4, "little", signed=True) + b"M\x00\x03\x86"
    item = b"B" + (2 * d * 4).to_bytes(4, "little")              # BINBYTES u32 length
    assert raw.count(item) == 1
    head = raw[:raw.index(item)].replace(shape_ops, big_shape)
    payload_bytes = n * d * 4
    assert payload_bytes > 1 << 32
    head += b"\x8e" + payload_bytes.to_bytes(8, "little")        # BINBYTES8 u64 length
    row = np.arange(d, dtype=np.float32) + 0.5
    with tempfile.TemporaryDirectory() as tmp:
        p = os.path.join(tmp, "passage_emb_block_0.pb")
        with open(p, "wb") as h:
            h.write(head)
            h.truncate(len(head) + payload_bytes + 64)           # sparse: only the touched pages exist
        fd = os.open(p, os.O_RDWR)
        try:
            os.pwrite(fd, row.tobytes(), len(head) + (n - 1) * d * 4)     # the last row, beyond 4 GiB
        finally:
            os.close(fd)
        hdr = loader.parse_ndarray_pickle_header(p)
        assert hdr.shape == (n, d) and hdr.dtype == np.dtype("<f4")
        assert hdr.payload_offset == len(head) and hdr.payload_bytes == payload_bytes
        mm = np.memmap(p, dtype="<f4", mode="r", offset=hdr.payload_offset, shape=(n, d))
        assert np.array_equal(mm[n - 1], row) and not mm[n - 2].any()
        del mm
        with open(p, "r+b") as h:                                 # a truncated block is refused, not half-loaded
            h.truncate(len(head) + payload_bytes - 1)
        with pytest.raises(ValueError):
            loader.parse_ndarray_pickle_header(p)


def test_stale_native_block_does_not_override_regenerated_pickles():
    rng = np.random.default_rng(3)
    a = rng.standard_normal((50, 64)).astype(np.float32)
    with tempfile.TemporaryDirectory() as d:
        write_blocks(d, [a])
        nat = loader.convert_block_to_native(d, 0)
        assert loader.find_block(d, 0)[0] == nat                  # current: preferred
        write_blocks(d, [a[:40]])                                 # the encoder ran again: other row count, newer
        t = os.path.getmtime(nat) + 10
        os.utime(os.path.join(d, "passage_emb_block_0.pb"), (t, t))
        with pytest.warns(UserWarning):
            found = loader.find_block(d, 0)
        assert found[0].endswith(".pb") and found[1]().shape == (40,)
        os.remove(os.path.join(d, "passage_emb_block_0.pb"))      # pickles gone: the native file is all there is
        os.remove(os.path.join(d, "passage_embid_block_0.pb"))
        assert loader.find_block(d, 0)[0] == nat


@pytest.mark.parametrize("name", GOLDEN_MERGE_CASES)
def test_reference_loop_mirror_matches_golden(name):
    """haconvdr_b200.retrieval.search_one_by_one_with_faiss driven with the oracle index reproduces
    the reference function's output rank for rank (golden made by the reference's own code)."""
    g = load_golden(name)
    with tempfile.TemporaryDirectory() as d:
        write_blocks(d, g["blocks"], g["id_start"])
        nb = int(g.get("block_num", len(g["blocks"]) + 3))
        D, I = retrieval.search_one_by_one_with_faiss(nb, d, FlatIP(g["q"].shape[1]), g["q"], g["k"])
    assert D.dtype == np.float64 and I.dtype == np.int64 and D.shape == g["D"].shape
    assert np.array_equal(I, g["I"])
    np.testing.assert_allclose(D, g["D"], rtol=1e-6, atol=0)


def test_rank_pids_and_trec_writer_match_reference_run():
    g = load_golden("trec_run_dedup_d64")
    ranked = retrieval.rank_pids(g["D"], g["I"], g["offset2pid"].tolist(), g["k"])
    assert ranked == offsets_to_ranked_pids(g["D"], g["I"], g["offset2pid"].tolist(), g["k"])
    with tempfile.TemporaryDirectory() as d:
        p = retrieval.write_trec_run(os.path.join(d, "run.trec"), g["qids"].tolist(), ranked, g["k"])
        assert open(p).read() == str(g["run_text"])
    assert "".join(trec_lines(g["qids"].tolist(), ranked, g["k"])) == str(g["run_text"])


def test_prj_run_file_without_score_column_matches_reference():
    """The PRJ drivers (`src/test_PRJ_topiocqa.py:218-299`) write the run file without the score column."""
    g = load_golden("prj_run_d64")
    ranked = retrieval.rank_pids(g["D"], g["I"], g["offset2pid"].tolist(), g["k"])
    with tempfile.TemporaryDirectory() as d:
        p = retrieval.write_trec_run(os.path.join(d, "run.trec"), g["qids"].tolist(), ranked, g["k"], with_score=False)
        assert open(p).read() == str(g["run_text"])


def test_faiss_compat_surface_builds_without_gpu():
    from haconvdr_b200 import faiss_compat as faiss
    res = faiss.StandardGpuResources()
    res.setTempMemory(0)
    co = faiss.GpuMultipleClonerOptions()
    co.shard, co.usePrecomputed = True, False
    vres, vdev = faiss.GpuResourcesVector(), faiss.Int32Vector()
    vdev.push_back(0)
    vres.push_back(res)
    idx = faiss.index_cpu_to_gpu_multiple(vres, vdev, faiss.IndexFlatIP(768), co)   # lazy: no device touched
    assert idx.d == 768 and idx.ntotal == 0


def test_native_block_file_round_trip_and_conversion():
    """Engine-native block files: header, 4 KiB alignment, contiguous ids stored as a range,
    arbitrary ids as an array; converting the reference's pickle pair gives the same bytes as writing directly."""
    rng = np.random.default_rng(12)
    x = rng.standard_normal((1037, 768)).astype(np.float32)
    with tempfile.TemporaryDirectory() as d:
        p = loader.write_native_block(loader.native_block_path(d, 0), x, np.arange(500, 1537, dtype=np.int64))
        h = loader.read_native_header(p)
        assert (h.n_rows, h.d, h.ids_kind, h.id0, h.emb_offset) == (1037, 768, 1, 500, 4096)
        assert os.path.getsize(p) % 4096 == 0
        assert np.array_equal(loader.load_block_array(p), x)
        assert np.array_equal(loader.load_native_embid(p), np.arange(500, 1537))
        ids = rng.permutation(5000)[:1037].astype(np.int64)
        p1 = loader.write_native_block(loader.native_block_path(d, 1), x, ids)
        h1 = loader.read_native_header(p1)
        assert h1.ids_kind == 0 and h1.ids_offset % 4096 == 0
        assert np.array_equal(loader.load_native_embid(p1), ids)
        # pickle pair -> native, streamed
        write_blocks(d, [x, x[:10]], 500)
        os.remove(p)
        assert loader.convert_block_to_native(d, 0) == p
        assert np.array_equal(loader.load_block_array(p), x) and loader.read_native_header(p).id0 == 500
        with open(p, "r+b") as f:           # corrupt magic -> rejected
            f.write(b"XXXX")
        with pytest.raises(ValueError):
            loader.read_native_header(p)


@pytest.mark.parametrize("name", ["merge_3blocks_d64", "short_two_blocks_d64", "block_num_limit_d64"])
def test_reference_loop_over_native_blocks_matches_golden(name):
    """Same golden outputs when every block exists only as a native file (the pickles are deleted)."""
    g = load_golden(name)
    with tempfile.TemporaryDirectory() as d:
        write_blocks(d, g["blocks"], g["id_start"])
        for b in range(len(g["blocks"])):
            loader.convert_block_to_native(d, b)
            for p in loader.block_paths(d, b):
                os.remove(p)
        nb = int(g.get("block_num", len(g["blocks"]) + 3))
        D, I = retrieval.search_one_by_one_with_faiss(nb, d, FlatIP(g["q"].shape[1]), g["q"], g["k"])
    assert np.array_equal(I, g["I"])
    np.testing.assert_allclose(D, g["D"], rtol=1e-6, atol=0)


@pytest.mark.parametrize("d,n", [(64, 1), (64, 1000), (192, 77), (1024, 5)])
def test_native_block_file_other_dimensions(d, n):
    rng = np.random.default_rng(d + n)
    x = rng.standard_normal((n, d)).astype(np.float32)
    ids = (np.arange(n, dtype=np.int64) * 3 + 7)              # not a contiguous range: stored as an array
    with tempfile.TemporaryDirectory() as tmp:
        p = loader.write_native_block(loader.native_block_path(tmp, 4), x, ids)
        h = loader.read_native_header(p)
        assert (h.n_rows, h.d) == (n, d) and h.emb_offset == 4096 and os.path.getsize(p) % 4096 == 0
        assert h.ids_kind == (1 if n == 1 else 0)             # a single id is trivially a range
        assert np.array_equal(loader.load_block_array(p), x)
        assert np.array_equal(loader.load_native_embid(p), ids)
        found = loader.find_block(tmp, 4)
        assert found is not None and found[0] == p and np.array_equal(found[1](), ids)
        assert loader.find_block(tmp, 5) is None
