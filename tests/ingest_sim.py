"""Synthetic paired-end alignments around one STR, written as SAM text (test infrastructure for tests/test_ingest.py and
tests/golden/make_bam_fixture.py).  The scenarios are built to reach every branch of read_and_filter_reads: low-quality
ends, adapter read-through, soft / hard clips, N bases, indels near the read ends, repeats that break the end-match
filters, XA / SA / AS / XS tags, mates far away or missing, both mates over the STR, duplicates, several read groups,
samples, libraries and files."""
import numpy as np

ADAPTERS = {"r1": "AGATCGGAAGAGCAC", "r2": "AGATCGGAAGAGCGT", "nx": "CTGTCTCTTATACAC"}
COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}


def revcomp(s):
    return "".join(COMP[c] for c in reversed(s))


def rand_seq(rng, n):
    return "".join("ACGT"[int(x)] for x in rng.integers(0, 4, n))


class Scenario:
    def __init__(self, seed, n_files=2, n_fragments=160, name_suffix=False, groups=None, tag="", rgs_per_file=None, hp_tags=False):
        rng = self.rng = np.random.default_rng(seed)
        self.hp_tags = hp_tags
        self.tag = tag
        period = int(rng.integers(2, 5))
        motif = rand_seq(rng, period)
        while len(set(motif)) == 1:
            motif = rand_seq(rng, period)
        copies = int(rng.integers(8, 20))
        left, right = rand_seq(rng, 4000), rand_seq(rng, 4000)
        if rng.random() < 0.5:      # a second, shorter copy of the repeat close by
            right = right[:40] + motif * int(rng.integers(3, 8)) + right[40:]
        self.chrom = left + motif * copies + right
        self.region = (4000, 4000 + period * copies)
        self.period = period
        self.name_suffix = name_suffix
        self.files = []
        for f in range(n_files):
            own = [("f%dg%d" % (f, g), "S%d" % int(rng.integers(0, 3)), "L%d" % int(rng.integers(0, 2))) for g in range(int(rng.integers(1, 4)))]
            if rgs_per_file:      # many samples: one per read group
                own = [("f%dg%d" % (f, g), "S%03d" % (f * rgs_per_file + g), "L%d" % (g % 2)) for g in range(rgs_per_file)]
            self.files.append({"groups": groups[f] if groups else own, "records": []})
        for i in range(n_fragments):
            self.fragment(i)

    # ---- one read ---------------------------------------------------------------------------
    def make_read(self, pos, length, reverse, which):
        rng = self.rng
        ops, bases = [], []
        ref = self.chrom
        at = pos
        err = 0.05 if rng.random() < 0.1 else 0.004
        indel_at = None
        if at < self.region[1] and at + length > self.region[0] and rng.random() < 0.35:
            indel_at = int(rng.integers(max(self.region[0], at + 1), max(self.region[0], at + 1) + 1 + max(0, min(self.region[1], at + length - 2) - max(self.region[0], at + 1))))
        if rng.random() < 0.06:       # an indel a few bases from a read end
            indel_at = at + int(rng.integers(2, 9)) if rng.random() < 0.5 else at + length - int(rng.integers(3, 10))
        n = 0
        run = 0
        while n < length:
            if indel_at is not None and at == indel_at and run > 0:
                k = self.period * int(rng.integers(1, 3))
                ops.append(("M", run)); run = 0
                if rng.random() < 0.5:
                    ops.append(("D", k)); at += k
                else:
                    ops.append(("I", k)); bases.append(rand_seq(rng, k)); n += k
                indel_at = None
                continue
            b = ref[at]
            if rng.random() < err:
                b = "ACGT"[("ACGT".index(b) + int(rng.integers(1, 4))) % 4]
            bases.append(b); at += 1; n += 1; run += 1
        if run:
            ops.append(("M", run))
        if ops[-1][0] != "M":       # never end on an indel
            ops.append(("M", 1)); bases.append(ref[at]); at += 1
        seq = "".join(bases)
        # clips
        if rng.random() < 0.12:
            k = int(rng.integers(1, 20)); seq = rand_seq(rng, k) + seq; ops.insert(0, ("S", k))
        if rng.random() < 0.12:
            k = int(rng.integers(1, 20)); seq = seq + rand_seq(rng, k); ops.append(("S", k))
        if rng.random() < 0.03:
            ops.insert(0, ("H", int(rng.integers(1, 30))))
        if rng.random() < 0.03:
            ops.append(("H", int(rng.integers(1, 30))))
        # adapter read-through: the 3' end of the ORIGINAL read, i.e. the start of a reverse-strand alignment
        if rng.random() < 0.15:
            ad = ADAPTERS[str(rng.choice(["r1", "nx"] if which == 1 else ["r2", "nx"]))]
            k = int(rng.integers(4, 31))
            tail = (ad + rand_seq(rng, 40))[:k]
            if rng.random() < 0.4 and k > 8:      # one mismatch inside the adapter
                j = int(rng.integers(0, min(k, len(ad))))
                tail = tail[:j] + "ACGT"[("ACGT".index(tail[j]) + 1) % 4] + tail[j + 1:]
            if k < len(seq):
                seq = (revcomp(tail) + seq[k:]) if reverse else (seq[:-k] + tail)
        if rng.random() < 0.03:
            j = int(rng.integers(0, len(seq))); seq = seq[:j] + "N" + seq[j + 1:]
        q = rng.integers(25, 41, len(seq))
        if rng.random() < 0.35:       # low-quality ends
            a, b = min(int(rng.integers(0, 30)), len(seq)), min(int(rng.integers(0, 30)), len(seq))
            q[:a] = rng.integers(2, 21, a)
            if b:
                q[-b:] = rng.integers(2, 21, b)
        if rng.random() < 0.04:
            q[:] = rng.integers(2, 12, len(seq))
        if rng.random() < 0.02:
            q[:] = 2
        quals = "".join(chr(33 + int(x)) for x in q)
        return seq, quals, ops

    def tags(self, rg, ops):
        rng = self.rng
        t = ["RG:Z:" + rg]
        if getattr(self, "hp", None):          # 10X haplotype tag of the fragment; a few reads lack it or disagree with their mate
            r = rng.random()
            if r < 0.85:
                t.append("HP:i:%d" % self.hp)
            elif r < 0.92:
                t.append("HP:i:%d" % (3 - self.hp))
        if rng.random() < 0.5:
            a = int(rng.integers(60, 150)); t += ["AS:i:%d" % a, "XS:i:%d" % max(0, a - int(rng.integers(0, 40)))]
        r = rng.random()
        cig = "".join("%d%s" % (n, c) for c, n in ops)
        if r < 0.06:
            t.append("XA:Z:chr2,+%d,100M,2;" % int(rng.integers(100, 4000)))
        elif r < 0.10:
            t.append("XA:Z:chr1,-%d,100M,1;" % int(rng.integers(3900, 4300)))
        elif r < 0.14:
            t.append("XA:Z:chr1_KI1_alt,+%d,%s,0;" % (int(rng.integers(100, 2000)), cig if rng.random() < 0.6 else "90M"))
        elif r < 0.17:
            t.append("XA:Z:chr2,+%d,100M,2;chr1,+%d,100M,3;" % (int(rng.integers(100, 4000)), int(rng.integers(100, 7000))))
        if rng.random() < 0.04:
            t.append("SA:Z:chr2,%d,+,50M50S,60,1;" % int(rng.integers(100, 4000)))
        return t

    def fragment(self, i):
        rng = self.rng
        f = int(rng.integers(0, len(self.files)))
        groups = self.files[f]["groups"]
        rg = groups[int(rng.integers(0, len(groups)))][0]
        L = int(rng.integers(90, 151))
        insert = int(rng.integers(L + 5, 650))
        start = int(rng.integers(self.region[0] - 700, self.region[1] + 200))
        name = "%sfrag%d" % (self.tag, i)
        if getattr(self, "hp_tags", False):
            self.hp = int(rng.integers(1, 3)) if rng.random() < 0.8 else None
        copies = 2 if rng.random() < 0.15 else 1          # PCR duplicates: same coordinates, another name
        for c in range(copies):
            nm = name if c == 0 else name + "dup"
            if rng.random() < 0.1:                          # single-end read
                rev = bool(rng.random() < 0.5)
                seq, quals, ops = self.make_read(start, L, rev, 1)
                self.add(f, nm, 16 if rev else 0, start, ops, "*", -1, seq, quals, self.tags(rg, ops))
                continue
            p1, p2 = start, start + insert - L
            if rng.random() < 0.05:
                p2 = start + int(rng.integers(1500, 3000))  # mate too far away
            s1, q1, o1 = self.make_read(p1, L, False, 1)
            s2, q2, o2 = self.make_read(p2, L, True, 2)
            f1, f2 = 0x1 | 0x2 | 0x20 | 0x40, 0x1 | 0x2 | 0x10 | 0x80
            if rng.random() < 0.5:                          # swap which mate is the forward one
                f1, f2 = 0x1 | 0x2 | 0x20 | 0x80, 0x1 | 0x2 | 0x10 | 0x40
            r = rng.random()
            if r < 0.02:
                f1 &= ~0xC0; f2 &= ~0xC0                    # paired, but neither first nor second
            elif r < 0.04:
                f2 = (f2 & ~0xC0) | (f1 & 0xC0)             # both claim to be the same end
            n1, n2 = (nm + "/1", nm + "/2") if self.name_suffix else (nm, nm)
            if f1 & 0x80 and self.name_suffix:
                n1, n2 = n2, n1
            drop_mate = rng.random() < 0.06
            self.add(f, n1, f1, p1, o1, "=", p2, s1, q1, self.tags(rg, o1))
            if not drop_mate:
                self.add(f, n2, f2, p2, o2, "=", p1, s2, q2, self.tags(rg, o2))
            elif rng.random() < 0.5:                        # unmapped mate placed at the read's position
                self.add(f, n2, (f2 | 0x4) & ~0x2, p1, [], "=", p1, rand_seq(rng, L), "I" * L, ["RG:Z:" + rg])

    def add(self, f, name, flag, pos, ops, rnext, pnext, seq, quals, tags):
        cigar = "".join("%d%s" % (n, c) for c, n in ops) or "*"
        fields = [name, str(flag), "chr1", pos + 1, "60", cigar, rnext, pnext + 1, "0", seq, quals] + tags
        self.files[f]["records"].append((pos, len(self.files[f]["records"]), fields))

    # ---- output ------------------------------------------------------------------------------
    def sam_text(self, f):
        head = ["@HD\tVN:1.5\tSO:coordinate", "@SQ\tSN:chr1\tLN:%d" % len(self.chrom), "@SQ\tSN:chr2\tLN:5000", "@SQ\tSN:chr1_KI1_alt\tLN:3000"]
        for g, sample, lib in self.files[f]["groups"]:
            head.append("@RG\tID:%s\tSM:%s\tLB:%s" % (g, sample, lib))
        lines = ["\t".join(str(x) for x in r[2]) for r in sorted(self.files[f]["records"], key=lambda r: r[:2])]
        return "\n".join(head + lines) + "\n"

    def rg_map(self, paths):
        return {paths[f] + g: (sample, lib) for f in range(len(self.files)) for g, sample, lib in self.files[f]["groups"]}


class MultiScenario:
    """Several STRs on one chromosome: independent Scenarios laid end to end (same files and read groups)."""

    def __init__(self, seed, n_regions=4, n_files=2, n_fragments=220, rgs_per_file=None, repeat=1, hp_tags=False):
        self.parts = []
        groups = None
        for k in range(n_regions):
            part = Scenario(seed * 100 + k, n_files=n_files, n_fragments=n_fragments, groups=groups, tag="r%d_" % k, rgs_per_file=rgs_per_file, hp_tags=hp_tags)
            groups = [f["groups"] for f in part.files]
            self.parts.append(part)
        self.parts = self.parts * repeat       # the same reads again further along the chromosome (cheap large inputs)
        self.offsets = np.cumsum([0] + [len(p.chrom) for p in self.parts])
        self.chrom = "".join(p.chrom for p in self.parts)
        self.regions = [(int(o) + p.region[0], int(o) + p.region[1], p.period) for o, p in zip(self.offsets, self.parts)]
        self.files = self.parts[0].files

    def sam_text(self, f):
        head = ["@HD\tVN:1.5\tSO:coordinate", "@SQ\tSN:chr1\tLN:%d" % len(self.chrom), "@SQ\tSN:chr2\tLN:5000", "@SQ\tSN:chr1_KI1_alt\tLN:3000"]
        for g, sample, lib in self.files[f]["groups"]:
            head.append("@RG\tID:%s\tSM:%s\tLB:%s" % (g, sample, lib))
        lines = []
        for o, p in zip(self.offsets, self.parts):
            for pos, idx, fields in sorted(p.files[f]["records"], key=lambda r: r[:2]):
                x = list(fields)
                x[3], x[7] = x[3] + int(o), x[7] + int(o) if x[6] == "=" else x[7]
                lines.append("\t".join(str(v) for v in x))
        return "\n".join(head + lines) + "\n"

    def rg_map(self, paths):
        return self.parts[0].rg_map(paths)

    def fasta_text(self):
        rng = np.random.default_rng(5)
        out = []
        for name, seq in (("chr1", self.chrom), ("chr2", rand_seq(rng, 5000)), ("chr1_KI1_alt", rand_seq(rng, 3000))):
            out.append(">" + name)
            out += [seq[i:i + 60] for i in range(0, len(seq), 60)]
        return "\n".join(out) + "\n"

    def region_text(self, extra=()):
        rows = [("chr1", s + 1, e, p, (e - s) / p, "STR%d" % i) for i, (s, e, p) in enumerate(self.regions)] + list(extra)
        return "".join("%s\t%d\t%d\t%d\t%.1f\t%s\n" % r for r in rows)

    def snp_vcf_text(self, seed=1, density=120):
        """A phased SNP VCF over chr1 (+ one record on chr2): biallelic SNPs with 0|1, 1|0, 1|1, 0|0, 0/1, ./. and .|. calls, plus
        indels, multi-allelic sites and sites inside / next to the STRs, for the BAM samples except one and one foreign sample."""
        rng = np.random.default_rng(seed)
        bam_samples = sorted({s for f in self.files for _, s, _ in f["groups"]})
        samples = ["X9"] + (bam_samples[:-1] if len(bam_samples) > 1 else bam_samples)
        lines = ["##fileformat=VCFv4.2", "##contig=<ID=chr1,length=%d>" % len(self.chrom), "##contig=<ID=chr2,length=5000>",
                 '##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">', '##FORMAT=<ID=DP,Number=1,Type=Integer,Description="Depth">',
                 "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(samples)]
        calls = ["0|1", "1|0", "0|1", "1|0", "1|1", "0|0", "0/1", "./.", ".|.", "1/0"]
        positions = set()
        for start, stop, _ in self.regions:
            positions.update(int(x) for x in rng.integers(start - 1300, stop + 1300, (stop - start + 2600) // density))
            positions.update(range(start - 17, start - 12))          # around the skip padding
            positions.update(range(stop + 13, stop + 18))
        for pos in sorted(p for p in positions if 1 <= p <= len(self.chrom) - 2):
            ref = self.chrom[pos - 1]
            alt = "ACGT"[("ACGT".index(ref) + int(rng.integers(1, 4))) % 4]
            r = rng.random()
            if r < 0.06:
                ref, alt = self.chrom[pos - 1:pos + 1], ref                      # deletion
            elif r < 0.10:
                alt = alt + "," + "ACGT"[("ACGT".index(ref) + 2) % 4 if alt != "ACGT"[("ACGT".index(ref) + 2) % 4] else ("ACGT".index(ref) + 1) % 4]
            elif r < 0.13:
                alt = ref + "T"                                                   # insertion
            gts = [calls[int(rng.integers(0, len(calls)))] + ":%d" % int(rng.integers(1, 60)) for _ in samples]
            lines.append("chr1\t%d\t.\t%s\t%s\t.\tPASS\t.\tGT:DP\t%s" % (pos, ref, alt, "\t".join(gts)))
        lines.append("chr2\t100\t.\tA\tC\t.\tPASS\t.\tGT:DP\t" + "\t".join("0|1:9" for _ in samples))
        return "\n".join(lines) + "\n"
