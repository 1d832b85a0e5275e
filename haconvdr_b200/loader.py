"""Loader for the reference's embedding block pickles.

The corpus encoder writes, every ~2.5 M passages, ``pickle.dump(np.float32[n,768], protocol=4)``
to ``passage_emb_block_{i}.pb`` and ``pickle.dump(np.int64[n])`` to ``passage_embid_block_{i}.pb``
(`/root/reference/gen_doc_embeddings.py:127-155`); the search loop ``pickle.load``s both per
block (`/root/reference/src/test_HAConvDR_topiocqa.py:81-93`), which costs two full host copies
of 7.7 GB.  Here the pickle opcode stream is walked only up to the raw payload, and the payload
is ``readinto`` page-locked staging buffers (a reader thread keeps the next chunk in flight)
from which ``hac_add`` copies straight into the HBM-resident shard.
"""
from __future__ import annotations

import ctypes
import os
import pickle
import struct
import threading

import numpy as np

from . import _lib

# opcode -> number of fixed argument bytes (protocol <= 5 subset used by ndarray pickles)
_FIXED = {
    b"\x80": 1, b"\x95": 8, b"\x94": 0, b"\x93": 0, b"K": 1, b"M": 2, b"J": 4, b"\x85": 0, b"\x86": 0,
    b"\x87": 0, b"(": 0, b"t": 0, b")": 0, b"R": 0, b"b": 0, b"N": 0, b"\x88": 0, b"\x89": 0, b"q": 1,
    b"r": 4, b"h": 1, b"j": 4, b".": 0, b"G": 8, b"]": 0, b"}": 0, b"e": 0, b"a": 0, b"s": 0, b"u": 0,
    b"\x81": 0, b"\x92": 0, b"0": 0, b"2": 0,
}


class BlockHeader:
    def __init__(self, shape, dtype, payload_offset, payload_bytes):
        self.shape, self.dtype = tuple(shape), np.dtype(dtype)
        self.payload_offset, self.payload_bytes = payload_offset, payload_bytes


def parse_ndarray_pickle_header(path: str, max_header: int = 1 << 16) -> BlockHeader:
    """Locate the raw C-order payload of a pickled ndarray without reading it.

    Accepts protocol 2-5 pickles written by numpy 1.x (``numpy.core.multiarray``) and 2.x
    (``numpy._core.multiarray``).  Raises ValueError when the stream is not a plain in-band
    C-contiguous ndarray (callers then fall back to ``pickle.load``)."""
    with open(path, "rb") as f:
        head = f.read(max_header)
    pos, ints, strs, fortran = 0, [], [], None
    n = len(head)
    while pos < n:
        op = head[pos:pos + 1]
        pos += 1
        if op in (b"B", b"\x8e", b"C", b"\x96"):      # BINBYTES, BINBYTES8, SHORT_BINBYTES, BYTEARRAY8
            if op == b"C":
                ln, pos = head[pos], pos + 1
            elif op == b"B":
                ln, pos = struct.unpack_from("<I", head, pos)[0], pos + 4
            else:
                ln, pos = struct.unpack_from("<Q", head, pos)[0], pos + 8
            if ln <= 8:                                   # the b'b' placeholder of _reconstruct
                pos += ln
                continue
            payload_offset, payload_bytes = pos, ln
            break
        if op in (b"\x8c", b"U"):                       # SHORT_BINUNICODE / SHORT_BINSTRING
            ln = head[pos]
            strs.append(head[pos + 1:pos + 1 + ln].decode("latin-1"))
            pos += 1 + ln
        elif op in (b"X", b"T"):                        # BINUNICODE / BINSTRING
            ln = struct.unpack_from("<I", head, pos)[0]
            strs.append(head[pos + 4:pos + 4 + ln].decode("latin-1"))
            pos += 4 + ln
        elif op == b"c":                                # GLOBAL "module\nname\n"
            e1 = head.index(b"\n", pos)
            e2 = head.index(b"\n", e1 + 1)
            strs.append(head[pos:e1].decode("latin-1"))
            strs.append(head[e1 + 1:e2].decode("latin-1"))
            pos = e2 + 1
        elif op == b"\x8a":                             # LONG1
            ln = head[pos]
            ints.append(int.from_bytes(head[pos + 1:pos + 1 + ln], "little", signed=True))
            pos += 1 + ln
        elif op in _FIXED:
            ln = _FIXED[op]
            if op == b"K":
                ints.append(head[pos])
            elif op == b"M":
                ints.append(struct.unpack_from("<H", head, pos)[0])
            elif op == b"J":
                ints.append(struct.unpack_from("<i", head, pos)[0])
            elif op == b"\x88":
                fortran = True
            elif op == b"\x89":
                fortran = False
            pos += ln
        else:
            raise ValueError("unsupported pickle opcode %r at %d in %s" % (op, pos - 1, path))
    else:
        raise ValueError("no ndarray payload found in the first %d bytes of %s" % (max_header, path))
    if "_reconstruct" not in strs or "ndarray" not in strs:
        raise ValueError("%s is not a pickled ndarray" % path)
    if fortran:
        raise ValueError("%s holds a Fortran-ordered array" % path)
    dt = next((s for s in strs if len(s) in (2, 3) and s[0] in "fiu" and s[1:].isdigit()), None)
    if dt is None:
        raise ValueError("dtype not found in %s" % path)
    itemsize = int(dt[1:])
    # ints: [0 (placeholder shape), 1 (pickle version), shape..., 3 (dtype version), -1, -1, 0]
    shape = None
    if len(ints) >= 3 and ints[0] == 0 and ints[1] == 1:
        for ndim in (1, 2, 3, 4):
            cand = ints[2:2 + ndim]
            if len(cand) == ndim and all(c >= 0 for c in cand) and int(np.prod(cand)) * itemsize == payload_bytes:
                shape = cand
                break
    if shape is None:
        raise ValueError("shape not recoverable from %s" % path)
    if os.path.getsize(path) < payload_offset + payload_bytes:
        raise ValueError("%s is truncated" % path)
    return BlockHeader(shape, "<" + dt, payload_offset, payload_bytes)


# ------------------------------------------------------------------------------------------------
# Engine-native block file: what `gen_doc_embeddings.py:127-155` would write next to (or
# instead of) the two pickles of a block.  One file per block, little endian:
#
#   [0, 4096)            header: magic "HACBLK01", u32 version, u32 d, u64 n_rows, u64 emb_offset, u64 ids_offset,
#                        u32 ids_kind (0 = int64 array at ids_offset, 1 = contiguous range from id0), i64 id0
#   [emb_offset, ...)    n_rows * d fp32, C order; emb_offset = 4096, the payload is zero-padded to a 4 KiB multiple
#   [ids_offset, ...)    n_rows int64 (ids_kind 0), 4 KiB aligned
#
# Everything sits on 4 KiB boundaries and a group of 4 rows of d = 768 is 3 pages, so chunks can be read with
# O_DIRECT straight into the page-locked staging buffers (no page-cache copy, no pickle walk).
NATIVE_MAGIC = b"HACBLK01"
NATIVE_ALIGN = 4096
_NATIVE_HDR = struct.Struct("<8sIIQQQIq")


def native_block_path(block_dir: str, block_id: int) -> str:
    return os.path.join(block_dir, "passage_emb_block_%d.hacb" % block_id)


def _pad(n: int) -> int:
    return (n + NATIVE_ALIGN - 1) // NATIVE_ALIGN * NATIVE_ALIGN


def write_native_block(path: str, emb, embid) -> str:
    """Write one block (``emb`` float32 ``[n, d]``, ``embid`` int64 ``[n]``) in the native format.  Contiguous ids
    (the non-distributed encoder's output, `/root/reference/src/utils.py:140-143`) are stored as a range."""
    emb = np.ascontiguousarray(emb, dtype=np.float32)
    embid = np.ascontiguousarray(embid, dtype=np.int64)
    assert emb.ndim == 2 and embid.shape == (emb.shape[0],)
    n, d = emb.shape
    is_range = n == 0 or bool(np.array_equal(embid, np.arange(embid[0], embid[0] + n, dtype=np.int64)))
    emb_off = NATIVE_ALIGN
    ids_off = emb_off + _pad(n * d * 4)
    hdr = _NATIVE_HDR.pack(NATIVE_MAGIC, 1, d, n, emb_off, 0 if is_range else ids_off, 1 if is_range else 0,
                           int(embid[0]) if n else 0)
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(hdr.ljust(NATIVE_ALIGN, b"\0"))
        f.write(memoryview(emb).cast("B"))
        f.write(b"\0" * (_pad(n * d * 4) - n * d * 4))
        if not is_range:
            f.write(memoryview(embid).cast("B"))
            f.write(b"\0" * (_pad(n * 8) - n * 8))
    os.replace(tmp, path)
    return path


class NativeHeader:
    def __init__(self, d, n_rows, emb_offset, ids_offset, ids_kind, id0):
        self.d, self.n_rows, self.emb_offset = d, n_rows, emb_offset
        self.ids_offset, self.ids_kind, self.id0 = ids_offset, ids_kind, id0
        self.shape, self.dtype = (n_rows, d), np.dtype("<f4")
        self.payload_offset, self.payload_bytes = emb_offset, n_rows * d * 4


def read_native_header(path: str) -> NativeHeader:
    with open(path, "rb") as f:
        raw = f.read(_NATIVE_HDR.size)
    if len(raw) < _NATIVE_HDR.size:
        raise ValueError("%s is too short for a native block" % path)
    magic, version, d, n, emb_off, ids_off, kind, id0 = _NATIVE_HDR.unpack(raw)
    if magic != NATIVE_MAGIC or version != 1:
        raise ValueError("%s is not a version-1 native block" % path)
    need = emb_off + n * d * 4 if kind == 1 else ids_off + n * 8
    if emb_off % NATIVE_ALIGN or (kind == 0 and ids_off % NATIVE_ALIGN) or kind not in (0, 1) or os.path.getsize(path) < need:
        raise ValueError("%s has a corrupt or truncated native header" % path)
    return NativeHeader(d, n, emb_off, ids_off, kind, id0)


def load_native_embid(path: str, hdr: NativeHeader = None) -> np.ndarray:
    hdr = hdr or read_native_header(path)
    if hdr.ids_kind == 1:
        return np.arange(hdr.id0, hdr.id0 + hdr.n_rows, dtype=np.int64)
    with open(path, "rb") as f:
        f.seek(hdr.ids_offset)
        return np.fromfile(f, dtype="<i8", count=hdr.n_rows)


def convert_block_to_native(block_dir: str, block_id: int, chunk_bytes: int = 64 << 20) -> str:
    """Pickle pair -> native file, streaming the payload (never holds the 7.7 GB array in memory)."""
    emb_path, embid_path = block_paths(block_dir, block_id)
    out = native_block_path(block_dir, block_id)
    try:
        hdr = parse_ndarray_pickle_header(emb_path)
        if hdr.dtype != np.dtype("<f4") or len(hdr.shape) != 2:
            raise ValueError("unexpected layout")
    except ValueError:
        with open(emb_path, "rb") as h:
            return write_native_block(out, pickle.load(h), load_embid(embid_path))
    embid = load_embid(embid_path)
    n, d = hdr.shape
    assert embid.shape == (n,)
    is_range = n == 0 or bool(np.array_equal(embid, np.arange(embid[0], embid[0] + n, dtype=np.int64)))
    ids_off = NATIVE_ALIGN + _pad(n * d * 4)
    head = _NATIVE_HDR.pack(NATIVE_MAGIC, 1, d, n, NATIVE_ALIGN, 0 if is_range else ids_off, 1 if is_range else 0,
                            int(embid[0]) if n else 0)
    tmp = out + ".tmp"
    with open(emb_path, "rb") as src, open(tmp, "wb") as dst:
        dst.write(head.ljust(NATIVE_ALIGN, b"\0"))
        src.seek(hdr.payload_offset)
        left = hdr.payload_bytes
        while left:
            buf = src.read(min(chunk_bytes, left))
            if not buf:
                raise IOError("short read in %s" % emb_path)
            dst.write(buf)
            left -= len(buf)
        dst.write(b"\0" * (_pad(n * d * 4) - n * d * 4))
        if not is_range:
            dst.write(memoryview(embid).cast("B"))
            dst.write(b"\0" * (_pad(n * 8) - n * 8))
    os.replace(tmp, out)
    return out


def load_block_array(path: str) -> np.ndarray:
    """The whole embedding array of a block file on the host (what ``pickle.load`` gives the reference)."""
    if path.endswith(".hacb"):
        hdr = read_native_header(path)
        with open(path, "rb") as f:
            f.seek(hdr.emb_offset)
            return np.fromfile(f, dtype="<f4", count=hdr.n_rows * hdr.d).reshape(hdr.n_rows, hdr.d)
    with open(path, "rb") as h:
        return pickle.load(h)


def _native_is_current(nat: str, emb_path: str, embid_path: str) -> bool:
    """A native file is used only if it cannot be stale: not older than the pickles it was converted from (when
    they are still there), and with the row count of the current embedding pickle."""
    try:
        hdr = read_native_header(nat)
    except (ValueError, OSError):
        return False
    have = [p for p in (emb_path, embid_path) if os.path.isfile(p)]
    if any(os.path.getmtime(p) > os.path.getmtime(nat) for p in have):
        return False
    if os.path.isfile(emb_path):
        try:
            if parse_ndarray_pickle_header(emb_path).shape[0] != hdr.n_rows:
                return False
        except ValueError:
            pass                      # an exotic pickle: the mtime check above is all there is
    return True


def find_block(block_dir: str, block_id: int):
    """(embedding file, loader of its id array) of a block, preferring the native file when it is current (a native
    file left over from an earlier encoder run must not override regenerated pickles); None when the block is
    missing (the reference stops at the first missing block, `:94-95`)."""
    nat = native_block_path(block_dir, block_id)
    emb_path, embid_path = block_paths(block_dir, block_id)
    if os.path.isfile(nat):
        if _native_is_current(nat, emb_path, embid_path):
            return nat, (lambda: load_native_embid(nat))
        import warnings
        warnings.warn("%s is stale or unreadable (older than the block pickles, or another row count): ignored" % nat)
    if os.path.isfile(emb_path) and os.path.isfile(embid_path):
        return emb_path, (lambda: load_embid(embid_path))
    return None


class PinnedBuffer:
    def __init__(self, nbytes: int):
        self.ptr = ctypes.c_void_p()
        _lib.check(_lib.lib().hac_pinned_alloc(nbytes, ctypes.byref(self.ptr)), "hac_pinned_alloc")
        self.nbytes = nbytes
        self.view = (ctypes.c_uint8 * nbytes).from_address(self.ptr.value)

    def free(self):
        if self.ptr.value:
            _lib.lib().hac_pinned_free(self.ptr)
            self.ptr = ctypes.c_void_p()


def block_paths(block_dir: str, block_id: int):
    return (os.path.join(block_dir, "passage_emb_block_%d.pb" % block_id),
            os.path.join(block_dir, "passage_embid_block_%d.pb" % block_id))


def load_embid(path: str) -> np.ndarray:
    with open(path, "rb") as h:
        return np.ascontiguousarray(pickle.load(h), dtype=np.int64)


def stream_block_into(index, emb_path: str, chunk_bytes: int = 64 << 20, buffers=None, row_range=None,
                      direct=None, stats=None) -> int:
    """Append the rows of one embedding block file (pickle ``.pb`` or native ``.hacb``) to ``index`` (a
    FlatIPIndex) through pinned staging.  ``row_range=(lo, hi)`` appends only rows [lo, hi) of the block (a
    rank's slice of a block that straddles two shards).  Returns the number of rows appended.  Falls back to
    ``pickle.load`` + ``add`` for pickles the header walker does not understand.  ``direct=True`` (or
    ``HAC_LOADER_DIRECT=1``) reads native blocks with O_DIRECT when the file system allows it; ``stats`` (a dict)
    receives ``{"direct": bool}``."""
    stats = stats if stats is not None else {}
    native = emb_path.endswith(".hacb")
    try:
        hdr = read_native_header(emb_path) if native else parse_ndarray_pickle_header(emb_path)
        if hdr.dtype != np.dtype("<f4") or len(hdr.shape) != 2 or hdr.shape[1] != index.d:
            raise ValueError("unexpected block layout %s %s" % (hdr.dtype, hdr.shape))
    except ValueError:
        if native:
            raise
        with open(emb_path, "rb") as h:
            arr = pickle.load(h)
        if row_range is not None:
            arr = arr[row_range[0]:row_range[1]]
        index.add(arr)
        return int(arr.shape[0])
    row_bytes = hdr.shape[1] * 4
    # chunks start on a multiple of `align_rows` rows, i.e. on a 4 KiB file boundary of a native block
    align_rows = NATIVE_ALIGN // np.gcd(row_bytes, NATIVE_ALIGN)
    rows_per_chunk = max(align_rows, chunk_bytes // row_bytes // align_rows * align_rows)
    own = buffers is None
    bufs = buffers or [PinnedBuffer(rows_per_chunk * row_bytes + NATIVE_ALIGN) for _ in range(2)]
    rows_per_chunk = min(rows_per_chunk, max(1, (bufs[0].nbytes - NATIVE_ALIGN) // row_bytes // align_rows * align_rows))
    if direct is None:
        direct = os.environ.get("HAC_LOADER_DIRECT", "0") == "1"
    direct = bool(direct and native and hasattr(os, "O_DIRECT") and bufs[0].ptr.value % NATIVE_ALIGN == 0)
    L = _lib.lib()
    lo, hi = (0, hdr.shape[0]) if row_range is None else (max(0, row_range[0]), min(hdr.shape[0], row_range[1]))
    n_rows = max(0, hi - lo)
    if n_rows == 0:
        return 0
    chunks = [(r, min(rows_per_chunk, hi - r)) for r in range(lo, hi, rows_per_chunk)]
    filled = [threading.Semaphore(0) for _ in bufs]
    freed = [threading.Semaphore(1) for _ in bufs]
    err = []
    stop = threading.Event()

    n_readers = max(1, min(8, (os.cpu_count() or 2) // 2))

    def read_slice(fd, mv, file_off):
        got = 0
        while got < len(mv):
            r = os.preadv(fd, [mv[got:]], file_off + got)      # releases the GIL; slices read in parallel
            if not r:
                raise IOError("short read in %s" % emb_path)
            got += r

    def reader():
        try:
            from concurrent.futures import ThreadPoolExecutor
            fd = -1
            use_direct = direct and lo % align_rows == 0
            if use_direct:
                try:
                    fd = os.open(emb_path, os.O_RDONLY | os.O_DIRECT)
                except OSError:             # tmpfs and some overlay mounts refuse O_DIRECT
                    use_direct = False
            if fd < 0:
                fd = os.open(emb_path, os.O_RDONLY)
            stats["direct"] = use_direct
            try:
                with ThreadPoolExecutor(n_readers) as pool:
                    for i, (r0, nr) in enumerate(chunks):
                        b = i % len(bufs)
                        freed[b].acquire()
                        if stop.is_set():                      # the consumer failed: nothing left to read for
                            return
                        # O_DIRECT wants whole pages: the tail chunk reads into the file's zero padding
                        want = _pad(nr * row_bytes) if use_direct else nr * row_bytes
                        mv = memoryview(bufs[b].view).cast("B")[:want]
                        base = hdr.payload_offset + r0 * row_bytes
                        step = (len(mv) + n_readers - 1) // n_readers
                        step = (step + 4095) // 4096 * 4096
                        futs = [pool.submit(read_slice, fd, mv[o:o + step], base + o) for o in range(0, len(mv), step)]
                        for f in futs:
                            f.result()
                        filled[b].release()
            finally:
                os.close(fd)
        except Exception as e:   # surfaced on the consumer side
            err.append(e)
            for s in filled:
                s.release()

    t = threading.Thread(target=reader, daemon=True)
    t.start()
    try:
        for i, (_, nr) in enumerate(chunks):
            b = i % len(bufs)
            filled[b].acquire()
            if err:
                raise err[0]
            _lib.check(L.hac_add(index._h, nr, bufs[b].ptr), "hac_add")   # returns after the copy completed
            freed[b].release()
    finally:
        stop.set()                         # on a consumer-side error the reader is parked on `freed`: wake it up
        for sem in freed:
            sem.release()
        t.join()
        if own:
            for b in bufs:
                b.free()
    return n_rows
