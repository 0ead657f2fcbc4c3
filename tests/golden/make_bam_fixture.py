"""Writes the committed BAM fixture of tests/test_ingest.py::test_golden_bam_fixture.

Run in the build container (needs oracle/_ref/libhipstr_ref.so, i.e. the reference checkout):
    python tests/golden/make_bam_fixture.py
The BAM and its index are written by htslib, and the two expected outputs by the UNMODIFIED reference
(BamCramMultiReader, BamProcessor::read_and_filter_reads + remove_pcr_duplicates) through oracle/ref_bam_harness.cpp."""
import json
import os
import pathlib
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import test_ingest as T  # noqa: E402
from ingest_sim import Scenario  # noqa: E402

sc = Scenario(2024, n_files=1, n_fragments=150)
with tempfile.TemporaryDirectory() as tmp:
    paths = T.write_bams(sc, pathlib.Path(tmp))
    bam = os.path.join(HERE, "ingest_f0.bam")
    for ext in ("", ".bai"):
        with open(paths[0] + ext, "rb") as src, open(bam + ext, "wb") as dst:
            dst.write(src.read())
region = T.ref_region_reads([bam], "chr1", 3000, 5100).replace(bam, "BAM")
with open(os.path.join(HERE, "ingest_f0.region.txt"), "w") as fh:
    fh.write(region)
rg_map = sc.rg_map([bam])
with open(os.path.join(HERE, "ingest_f0.filtered.txt"), "w") as fh:
    fh.write(T.ref_filter([bam], sc, rg_map, T.DEFAULTS))
with open(os.path.join(HERE, "ingest_f0.meta.json"), "w") as fh:
    json.dump({"chrom": sc.chrom, "region": list(sc.region), "period": sc.period,
               "groups": {g: [s, l] for g, s, l in sc.files[0]["groups"]}}, fh)
print("wrote", bam, len(region.splitlines()), "records in the region dump")
