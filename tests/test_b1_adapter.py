"""Seam B1 as a maintainer would bind it: oracle/ref_b1_adapter.cpp -- a class with the constructor / genotype() /
write_vcf_record() signatures of the reference's SeqStutterGenotyper, COMPILED against the reference headers, taking the
reference's std::vector<Alignment> / RegionGroup / StutterModel* and calling the product through the C-ABI -- driven by
the caller's own sequence (genotyper_bam_processor.cpp:229-246) and writing through the reference's VCFWriter.  The record
must equal the one the unmodified reference class writes for the same reads.
GPU: against hipstr_b200/libhipstr_b200.so.  CPU (`not gpu`): the same adapter over the host simulation."""
import gzip
import json
import os
import subprocess
import sys

import pytest

import checkers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ADAPTER = os.path.join(ROOT, "oracle", "_ref", "libhipstr_b1_adapter.so")
needs = pytest.mark.skipif(checkers.ref() is None or not os.path.exists(ADAPTER), reason="oracle/_ref not built")

DRIVER = r'''
import ctypes as C, gzip, json, os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import numpy as np
C.CDLL(%(lib)r, mode=C.RTLD_GLOBAL)          # the hipstr_* symbols the adapter binds
ad = C.CDLL(%(adapter)r)
from hipstr_b200.capi import Synth, c_f64p, c_i32p, c_u8p, ptr
from ref_genotyper import LocusReads, RefGenotyper
s = Synth(n_loci=3, n_samples=6, reads_per_sample=14, n_alleles=3, read_len=100, seed=%(seed)d, stutter_rate=0.25, flank_snp_freq=0.3,
          mate_rate=0.3)
out = []
for l in range(s.n_loci):
    rd = LocusReads(s, l)
    path = os.path.join(%(tmp)r, "locus%%d.vcf.gz" %% l)
    st6 = np.array([0.95, 0.05, 0.05, 0.95, 0.01, 0.01])
    ad.b1_adapter_run.restype = C.c_int32
    ad.b1_adapter_run.argtypes = [C.c_int32, C.c_int32, C.c_int32, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p, C.c_void_p, C.c_void_p, c_i32p,
                                  C.c_void_p, c_i32p, c_f64p, c_f64p, C.c_char_p, C.c_int32, C.c_int32, C.c_int32, c_f64p, C.c_int32,
                                  c_u8p, C.c_char_p]
    ok = ad.b1_adapter_run(0, rd.n_samples, rd.n_reads, ptr(rd.sample_label, c_i32p), ptr(rd.name_id, c_i32p), ptr(rd.start, c_i32p),
                           ptr(rd.stop, c_i32p), ptr(rd.seq_off, c_i32p), rd.bases.ctypes.data, rd.quals.ctypes.data,
                           ptr(rd.cigar_off, c_i32p), rd.cigar_type.ctypes.data, ptr(rd.cigar_len, c_i32p), ptr(rd.log_p1, c_f64p),
                           ptr(rd.log_p2, c_f64p), rd.chrom_seq, rd.region[0], rd.region[1], rd.period, ptr(st6, c_f64p), rd.haploid,
                           ptr(rd.rev_strand, c_u8p), path.encode())
    lines = [x for x in gzip.open(path, "rt").read().splitlines() if x]
    ref = RefGenotyper(rd, reassemble_flanks=True)
    good = bool(ref.initialized and ref.genotype(1000, 4, 0.01))
    out.append({"ok": int(ok), "lines": lines, "ref_ok": good, "ref": ref.vcf().rstrip("\n") if good else None})
    ref.close()
print("RESULT " + json.dumps(out))
'''


def run(lib, tmp_path, seed):
    code = DRIVER % dict(root=ROOT, lib=lib, adapter=ADAPTER, seed=seed, tmp=str(tmp_path))
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    res = json.loads([x for x in p.stdout.splitlines() if x.startswith("RESULT ")][-1][7:])
    norm = lambda t: t.replace(":-0.00:", ":0.00:")
    n_records = 0
    for r in res:
        assert r["ok"] == int(r["ref_ok"])
        if r["ref_ok"]:
            assert len(r["lines"]) == 1 and norm(r["lines"][0]) == norm(r["ref"])
            n_records += 1
        else:
            assert r["lines"] == []
    assert n_records > 0


@needs
@pytest.mark.parametrize("seed", [7, 19])
def test_b1_adapter_over_host_simulation(tmp_path, seed):
    checkers.build_hostsim()
    run(os.path.join(ROOT, "tests", "hostsim", "libhipstr_hostsim.so"), tmp_path, seed)


@needs
@pytest.mark.gpu
@pytest.mark.parametrize("seed", [7, 19, 23])
def test_b1_adapter_on_gpu(tmp_path, seed):
    run(os.path.join(ROOT, "hipstr_b200", "libhipstr_b200.so"), tmp_path, seed)
