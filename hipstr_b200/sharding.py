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
