"""Seam B5: hipstr::VCFWriter (the reference's VCFWriter, src/vcf_writer.{h,cpp}): reorder buffer semantics and BGZF output."""
import gzip
import heapq
import struct

import numpy as np
import pytest

from hipstr_b200 import capi


def reference_model(records, pad=50):
    """The reference's algorithm with std::push_heap/pop_heap replaced by heapq on (pos, arrival) -- positions are made
    distinct in the test so the heap's tie order does not matter."""
    out, heap, chrom = [], [], None
    for c, pos, text in records:
        if c != chrom:
            while heap:
                out.append(heapq.heappop(heap)[1])
            chrom = c
        else:
            while heap and heap[0][0] < pos - pad:
                out.append(heapq.heappop(heap)[1])
        heapq.heappush(heap, (pos, text))
    while heap:
        out.append(heapq.heappop(heap)[1])
    return out


def make_records(seed, n=400):
    rng = np.random.default_rng(seed)
    recs = []
    for chrom in ("chr1", "chr2", "chrX"):
        starts = np.sort(rng.choice(np.arange(1000, 200000, 7), n, replace=False))
        for s in starts:
            pos = int(s - rng.integers(0, 45))            # a record may precede its region start by the padding
            recs.append((chrom, pos, "%s\t%d\t.\tA\tAT\t.\t.\tEND=%d;X=%s" % (chrom, pos, pos + 20, "q" * int(rng.integers(0, 300)))))
    # make positions unique per chromosome
    seen, uniq = set(), []
    for c, p, t in recs:
        if (c, p) not in seen:
            seen.add((c, p))
            uniq.append((c, p, t))
    return uniq


def write(path, header, recs):
    lib = capi.load()
    w = lib.hipstr_vcf_writer_open(path.encode())
    assert w
    assert lib.hipstr_vcf_writer_header(w, header.encode()) == 0
    for c, p, t in recs:
        assert lib.hipstr_vcf_writer_add_record(w, c.encode(), p, t.encode()) == 0
    lib.hipstr_vcf_writer_close(w)


def test_plain_text_order_matches_reference_algorithm(tmp_path):
    recs = make_records(1)
    header = "##fileformat=VCFv4.1\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n"
    path = str(tmp_path / "out.vcf")
    write(path, header, recs)
    got = open(path).read()
    assert got.startswith(header)
    lines = got[len(header):].splitlines()
    assert lines == reference_model(recs)
    # within a chromosome the output is sorted by position even though the input was not
    for chrom in ("chr1", "chr2", "chrX"):
        pos = [int(l.split("\t")[1]) for l in lines if l.startswith(chrom + "\t")]
        assert pos == sorted(pos) and len(pos) > 100
    assert any(recs[i][1] > recs[i + 1][1] and recs[i][0] == recs[i + 1][0] for i in range(len(recs) - 1))


def test_bgzf_output_is_valid_blocked_gzip(tmp_path):
    recs = make_records(2, n=900)             # > 64 KiB so that several blocks are written
    header = "##fileformat=VCFv4.1\n"
    path = str(tmp_path / "out.vcf.gz")
    write(path, header, recs)
    raw = open(path, "rb").read()
    text = gzip.decompress(raw).decode()       # concatenated gzip members
    assert text == header + "".join(l + "\n" for l in reference_model(recs))
    # walk the blocks: gzip magic, FEXTRA with the BC subfield, BSIZE consistent, ISIZE <= 64 KiB; last block = EOF marker
    at, n_blocks, sizes = 0, 0, []
    while at < len(raw):
        assert raw[at:at + 4] == b"\x1f\x8b\x08\x04"
        xlen = struct.unpack_from("<H", raw, at + 10)[0]
        assert xlen == 6 and raw[at + 12:at + 16] == b"BC\x02\x00"
        bsize = struct.unpack_from("<H", raw, at + 16)[0] + 1
        isize = struct.unpack_from("<I", raw, at + bsize - 4)[0]
        assert isize <= 0x10000
        sizes.append(isize)
        at += bsize
        n_blocks += 1
    assert at == len(raw) and n_blocks >= 3 and sizes[-1] == 0
    assert raw[-28:] == bytes([0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0, 0x1b, 0, 0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0])


def test_misuse_is_reported_not_fatal(tmp_path):
    lib = capi.load()
    assert not lib.hipstr_vcf_writer_open(str(tmp_path / "no_such_dir" / "x.vcf").encode())
    assert lib.hipstr_vcf_writer_add_record(None, b"chr1", 1, b"x") == 3


def test_writer_reports_io_failure():
    """A write that fails (here: /dev/full, the kernel's always-full device) is reported by the call that hits it or by
    finish(), instead of leaving a silently truncated file."""
    import os
    if not os.path.exists("/dev/full"):
        pytest.skip("no /dev/full")
    from hipstr_b200 import capi
    lib = capi.load()
    w = lib.hipstr_vcf_writer_open(b"/dev/full")
    assert w
    lib.hipstr_vcf_writer_header(w, b"##fileformat=VCFv4.2\n" * 4000)
    for k in range(2000):
        lib.hipstr_vcf_writer_add_record(w, b"chr1", 100 + 60 * k, b"chr1\t%d\t.\tA\tC" % (100 + 60 * k))
    assert lib.hipstr_vcf_writer_finish(w) != 0
