"""BAM files -> VCF records: hipstr_process_regions (one window) against the UNMODIFIED reference program minus option
parsing (GenotyperBamProcessor::process_regions, forked over the host cores by regions) on the same synthetic files.
usage: python tools/bam_to_vcf_time.py [distinct_regions] [repeats] [fragments_per_region]      (needs a GPU + oracle/_ref)
100 samples (4 BAM files x 25 read groups); the distinct regions are repeated along the chromosome to get a large window."""
import json
import os
import pathlib
import sys
import tempfile
import time
from multiprocessing import Pool

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_pipeline as T
from ingest_sim import MultiScenario

n_distinct = int(sys.argv[1]) if len(sys.argv) > 1 else 8
repeat = int(sys.argv[2]) if len(sys.argv) > 2 else 12
n_fragments = int(sys.argv[3]) if len(sys.argv) > 3 else 5000
n_files, n_regions = 4, n_distinct * repeat
canon = lambda t: t.replace(":-0.00:", ":0.00:")


def reference_share(args):
    paths, fasta, bed, out, snp = args
    t = time.perf_counter()
    _, records = T.run_reference(paths, fasta, bed, out, 0, min_total_reads=100, snp_vcf=snp)
    return records, time.perf_counter() - t


if __name__ == "__main__":
    t0 = time.perf_counter()
    sc = MultiScenario(11, n_regions=n_distinct, n_files=n_files, n_fragments=n_fragments, rgs_per_file=25, repeat=repeat)
    with tempfile.TemporaryDirectory() as tmp:
        tp = pathlib.Path(tmp)
        paths, fasta, bed = T.files_of(sc, tp)
        snp = T.write_snp_vcf(sc, tp, 11)[1]
        gen_s = time.perf_counter() - t0
        cores = len(os.sched_getaffinity(0))
        workers = min(cores, n_regions)
        lines = open(bed).read().splitlines()
        shares = []
        for w in range(workers):
            share = str(tp / ("regions_%d.bed" % w))
            with open(share, "w") as fh:
                fh.write("\n".join(lines[w::workers]) + "\n")
            shares.append((paths, fasta, share, str(tp / ("ref_%d.vcf" % w)), snp))
        t = time.perf_counter()
        with Pool(workers) as pool:
            parts = pool.map(reference_share, shares)
        ref_wall = time.perf_counter() - t
        ref_cpu = sum(p[1] for p in parts)
        want = sorted((r for p in parts for r in p[0]), key=lambda r: int(r.split("\t")[1]))

        from hipstr_b200 import capi, pipeline
        opt = pipeline.Options(snp_vcf=snp)
        chroms, regions = pipeline.read_fasta(fasta), pipeline.read_regions(bed)
        with capi.Context(0) as ctx:
            t = time.perf_counter()
            pipeline.process_regions(ctx, paths, chroms, regions, opt)              # first window: device buffers grow to their size
            cold = time.perf_counter() - t
            t = time.perf_counter()
            records, summary = pipeline.process_regions(ctx, paths, chroms, regions, opt)
            ours = time.perf_counter() - t
            trace_s = ctx.trace_seconds()
        # the same regions as 3 windows on 3 threads / contexts: host stages of one window overlap device stages of another
        from concurrent.futures import ThreadPoolExecutor
        n_win = 3
        ctxs = [capi.Context(0) for _ in range(n_win)]
        parts_of = [regions[w * len(regions) // n_win:(w + 1) * len(regions) // n_win] for w in range(n_win)]
        run_win = lambda w: pipeline.process_regions(ctxs[w], paths, chroms, parts_of[w], opt)
        with ThreadPoolExecutor(n_win) as ex:
            list(ex.map(run_win, range(n_win)))                                     # sizes the buffers of every context
            t = time.perf_counter()
            outs = list(ex.map(run_win, range(n_win)))
            piped = time.perf_counter() - t
        piped_records = [r for o in outs for r in o[0]]
        same_piped = [canon(r[2]) for r in piped_records] == [canon(w) for w in want]
        for c in ctxs:
            c.close()
        same = [canon(r[2]) for r in records] == [canon(w) for w in want]
    print(json.dumps({"regions": n_regions, "distinct_regions": n_distinct, "samples": 100, "files": n_files, "alignments_read": summary["alignments_read"], "reads_kept": summary["reads_kept"],
                      "records": len(records), "identical_to_reference": same, "ours_s": round(ours, 3), "ours_loci_per_s": n_regions / ours, "ours_first_window_s": round(cold, 3),
                      "ours_3_windows_s": round(piped, 3), "ours_3_windows_loci_per_s": n_regions / piped, "ours_3_windows_identical": same_piped,
                      "stage_seconds": {k: round(v, 3) for k, v in summary["seconds"].items()},
                      "genotyper_seconds": summary["genotyper_seconds"], "hmm_alignments": summary["alignments"], "traces": summary["traces"],
                      "rounds": summary["rounds"], "trace_call_seconds_incl_warmup": {k: round(v, 3) for k, v in trace_s.items()},
                      "reference_wall_s": round(ref_wall, 2), "reference_cpu_s": round(ref_cpu, 2), "reference_workers": workers,
                      "reference_loci_per_s": n_regions / ref_wall, "reference_loci_per_s_per_core": n_regions / ref_cpu,
                      "host_cores": cores, "phased_reads": summary["phased_reads"], "data_generation_s": round(gen_s, 1),
                      "what": "EM-trained stutter models, phased SNP VCF, flank assembly on; reference = process_regions forked by regions"}))
    assert same and same_piped
