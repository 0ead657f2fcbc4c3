"""The N>1 path of bench.py on CPU: world size 2 over gloo.  Loci shard across ranks with no data-path
collective; the single collective is the gather of per-locus genotype records to rank 0.  The per-rank
"compute" here is the CPU oracle (test infrastructure) -- what is under test is the sharding, the
gather and the rank-0 merge order, which are device independent."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_loci, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import checkers
    from hipstr_b200.capi import Synth
    from hipstr_b200.sharding import shard_bounds, merge_records
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s = Synth(n_loci=n_loci, n_samples=4, reads_per_sample=6, n_alleles=3, read_len=100, seed=77)   # same loci on every rank
    l0, l1 = shard_bounds(n_loci, rank, world)
    o = checkers.oracle()
    ll = np.zeros(s.n_out)
    import ctypes as C
    from hipstr_b200.capi import c_f64p, ptr
    o.oracle_align_loci(C.byref(s.batch), l0, l1, ptr(ll, c_f64p), None)
    # per-locus record = (global locus index, checksum of its LLs)
    rec = torch.tensor([[l, ll[s.locus_out_off[l]:s.locus_out_off[l + 1]].sum()] for l in range(l0, l1)], dtype=torch.float64).reshape(-1, 2)
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([rec.shape[0]], dtype=torch.int64))
    width = int(max(c.item() for c in counts))
    padded = torch.full((width, 2), -1.0, dtype=torch.float64)
    padded[:rec.shape[0]] = rec
    gathered = [torch.zeros_like(padded) for _ in range(world)] if rank == 0 else None
    dist.gather(padded, gathered, dst=0)
    if rank == 0:
        merged = merge_records([g[:int(c.item())].numpy() for g, c in zip(gathered, counts)])
        np.save(out_path, merged)
    dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    sys.path.insert(0, ROOT)
    from hipstr_b200.sharding import shard_bounds
    for n in (0, 1, 7, 8, 1000):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(x[1] - x[0] for x in b) - min(x[1] - x[0] for x in b) <= 1


@pytest.mark.timeout(300)
def test_two_rank_gloo_gather_matches_single_rank(tmp_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    n_loci = 5
    out = str(tmp_path / "merged.npy")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_loci, out), nprocs=2, join=True)
    merged = np.load(out)
    import checkers
    from hipstr_b200.capi import Synth
    s = Synth(n_loci=n_loci, n_samples=4, reads_per_sample=6, n_alleles=3, read_len=100, seed=77)
    ll = checkers.align(checkers.oracle(), "oracle_", s.batch, s.n_out)
    want = np.array([[l, ll[s.locus_out_off[l]:s.locus_out_off[l + 1]].sum()] for l in range(n_loci)])
    assert merged.shape == want.shape
    assert np.array_equal(merged[:, 0], want[:, 0])          # rank 0 feeds records in locus (position) order
    assert np.array_equal(merged[:, 1], want[:, 1])


def _records_of(l0, l1):
    """Fake per-locus VCF records with out-of-order positions inside the writer's 50-bp tolerance."""
    return [(l, "chr1" if l < 4 else "chr2", 1000 + 100 * l - (30 if l % 3 == 1 else 0), "chrX\t%d\tSTR%d\tpayload-%s" % (l, l, "x" * (l % 5)))
            for l in range(l0, l1)]


def _record_worker(rank, world, port, n_loci, out_path):
    sys.path.insert(0, ROOT)
    from hipstr_b200.capi import load
    from hipstr_b200.sharding import gather_vcf_records, shard_bounds, write_records
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    l0, l1 = shard_bounds(n_loci, rank, world)
    merged = gather_vcf_records(_records_of(l0, l1))
    if rank == 0:
        lib = load()
        w = lib.hipstr_vcf_writer_open(out_path.encode())
        lib.hipstr_vcf_writer_header(w, b"##header\n")
        write_records(merged, w, lib)
        lib.hipstr_vcf_writer_close(w)
    else:
        assert merged is None
    dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", [2, 3])
def test_vcf_records_gathered_to_rank0_in_locus_order(tmp_path, world):
    """Sharded loci -> gather_vcf_records -> hipstr::VCFWriter on rank 0 writes the same file a single rank would."""
    sys.path.insert(0, ROOT)
    from hipstr_b200.capi import load
    from hipstr_b200.sharding import pack_records, unpack_records, write_records
    n_loci = 8
    recs = _records_of(0, n_loci)
    assert unpack_records(pack_records(recs)) == recs
    single = str(tmp_path / "single.vcf")
    lib = load()
    w = lib.hipstr_vcf_writer_open(single.encode())
    lib.hipstr_vcf_writer_header(w, b"##header\n")
    write_records(recs, w, lib)
    lib.hipstr_vcf_writer_close(w)
    sharded = str(tmp_path / "sharded.vcf")
    mp.spawn(_record_worker, args=(world, _free_port(), n_loci, sharded), nprocs=world, join=True)
    assert open(sharded).read() == open(single).read()
    assert open(single).read().count("\n") == n_loci + 1


def _loop_worker(rank, world, port, n_loci, out_path):
    """bench.py's N>1 loop on CPU: every rank drives the C++ multi-GPU driver (over the host simulation), the ranks share
    ONE locus list through a counter in the rendezvous store, rank 0 gathers the VCF records over gloo."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from hipstr_b200 import capi
    capi._lib, capi.LIB_PATH = None, os.path.join(ROOT, "tests", "hostsim", "libhipstr_hostsim.so")
    from hipstr_b200.sharding import StoreDealer, gather_vcf_records
    import test_hostsim
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["HIPSTR_HOST_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s = test_hostsim.synth(n_loci, 91)   # the same list on every rank
    m = capi.MultiGenotyper(devices=(0,), pipelines=2)
    dealer = StoreDealer(dist.distributed_c10d._get_default_store(), "loop_test")
    ok, rec = m.genotype_synth(s, test_hostsim.vcf_loci(s), 2, next_window=dealer)
    windows = sum(m.stats()["windows_per_worker"])
    m.close()
    merged = gather_vcf_records([(l, "chrS", r[0], r[1]) for l, r in enumerate(rec) if r is not None])
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([windows], dtype=torch.int64))
    if rank == 0:
        import json
        json.dump({"records": merged, "windows": [int(c.item()) for c in counts]}, open(out_path, "w"))
    dist.destroy_process_group()


def test_shared_list_loop_two_ranks(tmp_path):
    """World size 2: windows are dealt dynamically from one shared list, every window exactly once, and the records rank 0
    holds after the gather equal the single-process result in locus order."""
    import json
    import subprocess
    import checkers
    checkers.build_hostsim()
    n_loci = 9
    out = str(tmp_path / "loop.json")
    mp.spawn(_loop_worker, args=(2, _free_port(), n_loci, out), nprocs=2, join=True)
    got = json.load(open(out))
    assert sum(got["windows"]) == 5
    # the single-process result, through the same host simulation
    code = ("import sys, json; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "from hipstr_b200 import capi\n"
            "capi.LIB_PATH = %r\n"
            "import test_hostsim\n"
            "s = test_hostsim.synth(%d, 91)\n"
            "ok, rec = test_hostsim.single_context_records(s)\n"
            "print(json.dumps([[l, 'chrS', r[0], r[1]] for l, r in enumerate(rec) if r is not None]))\n"
            % (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "hostsim", "libhipstr_hostsim.so"), n_loci))
    want = json.loads(subprocess.run([sys.executable, "-c", code], check=True, capture_output=True, text=True).stdout.splitlines()[-1])
    assert got["records"] == want and len(want) > 0
