"""FASTA input as `pangraph build` reads it, through the C-ABI (include/pgmm_b200.h, Part 6).  Mirrors
packages/pangraph/src/io/fasta.rs: FastaReader::from_paths(..).read_many(), FastaReader::from_str(..).read_many()."""
import ctypes as C
from dataclasses import dataclass
from typing import Optional

from . import abi

DNA = "ACGTYRWSKMDVHBN"           # Alphabet::DnaWithoutGap (io/fasta.rs:273-277), the reader's default
DNA_WITH_GAP = "ACGTYRWSKMDVHBN-"  # Alphabet::DnaWithGap


class FastaError(RuntimeError):
    pass


@dataclass
class FastaRecord:
    seq_name: str
    desc: Optional[str]
    seq: bytes
    index: int


class pgmm_fasta_record_t(C.Structure):
    _fields_ = [("name", C.c_void_p), ("desc", C.c_void_p), ("seq", C.c_void_p), ("len", C.c_int64), ("index", C.c_int64)]


def _collect(L, rc, recs, n, err):
    if rc != 0:
        raise FastaError(err.value.decode(errors="replace"))
    try:
        return [FastaRecord(C.string_at(recs[i].name).decode(), C.string_at(recs[i].desc).decode() if recs[i].desc else None,
                            C.string_at(recs[i].seq, recs[i].len), int(recs[i].index)) for i in range(n.value)]
    finally:
        L.pgmm_fasta_free.argtypes = [C.POINTER(pgmm_fasta_record_t), C.c_int64]
        L.pgmm_fasta_free.restype = None
        L.pgmm_fasta_free(recs, n.value)


def read_many_str(contents, alphabet=None):
    L = abi.lib()
    data = contents.encode() if isinstance(contents, str) else bytes(contents)
    recs, n, err = C.POINTER(pgmm_fasta_record_t)(), C.c_int64(0), C.create_string_buffer(2048)
    L.pgmm_fasta_read_buffer.restype = C.c_int
    rc = L.pgmm_fasta_read_buffer(data, C.c_int64(len(data)), alphabet.encode() if alphabet else None, C.byref(recs), C.byref(n), err, len(err))
    return _collect(L, rc, recs, n, err)


def read_many(paths, alphabet=None):
    L = abi.lib()
    paths = [paths] if isinstance(paths, (str, bytes)) else list(paths)
    arr = (C.c_char_p * max(1, len(paths)))(*[p.encode() if isinstance(p, str) else p for p in paths])
    recs, n, err = C.POINTER(pgmm_fasta_record_t)(), C.c_int64(0), C.create_string_buffer(2048)
    L.pgmm_fasta_read_files.restype = C.c_int
    rc = L.pgmm_fasta_read_files(len(paths), arr, alphabet.encode() if alphabet else None, C.byref(recs), C.byref(n), err, len(err))
    return _collect(L, rc, recs, n, err)
