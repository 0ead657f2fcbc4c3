"""Locus sharding across GPUs: loci are independent, so rank r owns a contiguous slice of the
sorted locus list, computes it with no communication, and rank 0 gathers the finished per-locus records and
merges them back into locus order -- the order VCFWriter::add_vcf_record needs (src/vcf_writer.h:33-35)."""
import numpy as np


def shard_bounds(n_loci, rank, world):
    """Contiguous, balanced (sizes differ by at most 1) slice [lo, hi) of rank `rank`."""
    base, extra = divmod(n_loci, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def merge_records(per_rank):
    """per_rank: list (by rank) of arrays whose first column is the global locus index.  Returns one array sorted
    by locus index (stable), i.e. what rank 0 feeds to the VCF writer."""
    rows = [r for r in per_rank if len(r)]
    if not rows:
        return np.zeros((0, 2))
    allr = np.concatenate(rows, axis=0)
    return allr[np.argsort(allr[:, 0], kind="stable")]


def pack_records(records):
    """records: [(global locus index, chrom, pos, text)] -> one uint8 array:
    per record int64 locus, int32 pos, int32 len(chrom), int32 len(text), then the two strings."""
    import struct
    out = bytearray()
    for locus, chrom, pos, text in records:
        c, t = chrom.encode(), text.encode()
        out += struct.pack("<qiii", locus, pos, len(c), len(t)) + c + t
    return np.frombuffer(bytes(out), dtype=np.uint8).copy()


def unpack_records(buf):
    import struct
    raw, at, out = bytes(buf), 0, []
    while at < len(raw):
        locus, pos, nc, nt = struct.unpack_from("<qiii", raw, at)
        at += 20
        out.append((locus, raw[at:at + nc].decode(), pos, raw[at + nc:at + nc + nt].decode()))
        at += nc + nt
    return out


def gather_vcf_records(records, device=None):
    """The one collective of the sharded path: every rank's finished VCF records -> rank 0, merged back
    into locus order.  Works on any initialised torch.distributed backend: byte tensors live on `device` (a CUDA device
    under NCCL, None = CPU under gloo).  Returns the merged list on rank 0, None elsewhere."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    mine = torch.from_numpy(pack_records(records))
    if device is not None:
        mine = mine.to(device)
    n = torch.tensor([mine.numel()], dtype=torch.int64, device=mine.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    width = max(int(x.item()) for x in sizes)
    padded = torch.zeros(max(width, 1), dtype=torch.uint8, device=mine.device)
    padded[:mine.numel()] = mine
    parts = [torch.zeros_like(padded) for _ in range(world)] if rank == 0 else None
    dist.gather(padded, parts, dst=0)
    if rank != 0:
        return None
    merged = []
    for part, size in zip(parts, sizes):
        merged.extend(unpack_records(part[:int(size.item())].cpu().numpy()))
    merged.sort(key=lambda r: r[0])   # stable: locus order == (chromosome, position) order of the sorted region list
    return merged


def write_records(merged, writer_handle, lib):
    """Feed merged records to a hipstr_vcf_writer_t in order (VCFWriter::add_vcf_record)."""
    for _, chrom, pos, text in merged:
        st = lib.hipstr_vcf_writer_add_record(writer_handle, chrom.encode(), pos, text.encode())
        if st != 0:
            raise RuntimeError("hipstr_vcf_writer_add_record failed with status %d" % st)


def lpt_assign(costs, world):
    """Cost-aware static sharding: longest-processing-time-first assignment of units (loci or windows)
    with the given costs to `world` ranks.  Returns rank_of[unit].  Deterministic, so every rank computes the same map."""
    import heapq
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    heap = [(0, r) for r in range(world)]
    rank_of = [0] * len(costs)
    for i in order:
        load, r = heapq.heappop(heap)
        rank_of[i] = r
        heapq.heappush(heap, (load + costs[i], r))
    return rank_of


class StoreDealer:
    """Dynamic window dealing across the ranks of a torch.distributed job: one atomic counter in the rendezvous store
    (TCPStore.add).  Passed as hipstr_multi_genotype's next_window callback, so a rank pulls the next window of the
    SHARED locus list the moment one of its pipelines is free -- no static split, no data-path collective."""

    def __init__(self, store, key):
        self.store, self.key = store, key

    def __call__(self):
        return int(self.store.add(self.key, 1)) - 1
