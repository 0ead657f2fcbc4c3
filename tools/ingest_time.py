"""Host ingestion (BAM region query -> read_and_filter_reads -> PCR-duplicate removal) of this repo against the UNMODIFIED
reference over htslib, on the same synthetic BAM files, one host core each.
usage: python tools/ingest_time.py [n_fragments] [n_regions]      (CPU only; needs oracle/_ref/libhipstr_ref.so)"""
import json
import os
import pathlib
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_ingest as T
from hipstr_b200 import capi
from ingest_sim import Scenario

n_fragments = int(sys.argv[1]) if len(sys.argv) > 1 else 3000      # ~100 samples x 30 reads around one STR
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
sc = Scenario(77, n_files=4, n_fragments=n_fragments)
with tempfile.TemporaryDirectory() as tmp:
    paths = T.write_bams(sc, pathlib.Path(tmp))
    rg_map = sc.rg_map(paths)
    want = T.ref_filter(paths, sc, rg_map, T.DEFAULTS)
    got, counts, _ = T.ours_filter(paths, sc, rg_map, T.DEFAULTS)
    assert got == want
    t = time.perf_counter()
    for _ in range(reps):
        T.ref_filter(paths, sc, rg_map, T.DEFAULTS)
    ref_s = (time.perf_counter() - t) / reps
    t = time.perf_counter()
    for _ in range(reps):
        reader = capi.BamReader(paths)
        recs = reader.fetch("chr1", sc.region[0] - 1000, sc.region[1] + 1000)
        filtered = recs.filter(sc.chrom, [sc.region], rg_map)
        n_records = len(recs)
    ours_s = (time.perf_counter() - t) / reps
    # the stages of the product separately
    reader = capi.BamReader(paths)
    t = time.perf_counter()
    for _ in range(reps):
        recs = reader.fetch("chr1", sc.region[0] - 1000, sc.region[1] + 1000)
    fetch_s = (time.perf_counter() - t) / reps
    t = time.perf_counter()
    for _ in range(reps):
        filtered = recs.filter(sc.chrom, [sc.region], rg_map)
    filter_s = (time.perf_counter() - t) / reps
print(json.dumps({"files": len(paths), "records_in_region": n_records, "reads_kept": counts["passed"], "identical_to_reference": True,
                  "reference_ms_per_region": round(ref_s * 1e3, 2), "ours_ms_per_region": round(ours_s * 1e3, 2),
                  "ours_fetch_ms": round(fetch_s * 1e3, 2), "ours_filter_ms": round(filter_s * 1e3, 2),
                  "records_per_s_ours": n_records / ours_s, "records_per_s_reference": n_records / ref_s,
                  "note": "both open the files and load the indexes every repetition; one core"}))
