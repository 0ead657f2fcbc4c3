/*
 * hipstr_synth.h -- deterministic synthetic STR loci for tests and bench.py
 * (SURVEY.md 8d).  Not part of the drop-in boundary: it only PRODUCES the flat
 * inputs that include/hipstr_b200.h consumes, the way
 * GenotyperBamProcessor::analyze_reads_and_phasing would after BAM parsing,
 * trimming and left-alignment (genotyper_bam_processor.cpp:161-243).
 */
#ifndef HIPSTR_SYNTH_H_
#define HIPSTR_SYNTH_H_
#include "../../include/hipstr_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int32_t n_loci;
  int32_t n_samples;
  int32_t reads_per_sample;
  int32_t n_alleles;        /* requested candidate STR alleles (ref +/- copies, >= 2 copies) */
  int32_t read_len;
  int32_t trim;             /* 1 = clip reads to STR +/- 40 bp like TrimAlignment */
  int32_t period;           /* motif length (default 4) */
  int32_t ref_copies;       /* reference copy number (default 12) */
  uint64_t seed;
  double stutter_rate;      /* P(read carries +/- 1 copy), default 0.05 */
  double sub_rate;          /* flank substitution rate, default 1/200 */
  double mate_rate;         /* P(read is followed by an adjacent second mate), default 0 */
  double flank_snp_freq;    /* population frequency of a planted SNP 15 bp upstream and one 12 bp downstream of
                               the STR (each sample chromosome draws them independently), default 0 */
  int32_t haploid;          /* 1 = haploid loci (one chromosome copy per sample, Genotyper's haploid_ flag), default 0 */
} hipstr_synth_cfg_t;

typedef struct {
  hipstr_align_batch_t batch;      /* pooled reads + haplotype blocks, ready for hipstr_align_batch_* */
  int64_t n_reads;                 /* un-pooled reads over all loci */
  const int32_t* locus_read_off;   /* [n_loci+1] */
  const int32_t* locus_sample_off; /* [n_loci+1] */
  const int32_t* pool_index;       /* [n_reads] pool of the read, local to its locus */
  const int32_t* sample_label;     /* [n_reads] sample of the read, local to its locus */
  const uint8_t* second_mate;      /* [n_reads] */
  const int32_t* read_weight;      /* [n_reads] 0 for second mates (seq_stutter_genotyper.cpp:499-500) */
  const double*  log_p1;           /* [n_reads] */
  const double*  log_p2;           /* [n_reads] */
  const int32_t* n_haps;           /* [n_loci] */
  const uint8_t* haploid;          /* [n_loci] */
  const int32_t* true_gt;          /* [total samples][2] simulated allele indices */
  const int32_t* read_bp_diff;     /* [n_reads] bp difference of the read's STR vs the reference */
  int64_t read_ll_size;            /* sum over loci R_l * H_l */
  int64_t post_size;               /* sum over loci S_l * H_l^2 */
  /* the un-pooled reads as seam B1 receives them (std::vector<Alignment>, sample-major) */
  const int32_t* read_seq_off;     /* [n_reads+1] offsets into read_bases / read_quals */
  const char*    read_bases;
  const char*    read_quals;
  const int32_t* read_start;       /* [n_reads] Alignment::get_start() */
  const int32_t* read_cigar_off;   /* [n_reads+1] */
  const char*    read_cigar_type;  /* '=', 'X', 'I', 'D' */
  const int32_t* read_cigar_len;
  const int32_t* read_name_id;     /* [n_reads] equal ids on adjacent reads = mates (same read name) */
  const int32_t* block_start;      /* [n_blocks] HapBlock::start() */
  const int32_t* block_end;        /* [n_blocks] HapBlock::end() */
  int32_t chrom_len;               /* every locus has its own chromosome of this length */
  const char* chrom_seqs;          /* [n_loci][chrom_len] */
  int32_t region_start, region_stop; /* the STR Region of every locus */
  const int32_t* read_stop;        /* [n_reads] Alignment::get_stop(): last aligned reference position (inclusive) */
  const uint8_t* read_rev_strand;  /* [n_reads] Alignment::is_from_reverse_strand(): a hash of the read index */
} hipstr_synth_view_t;

typedef struct hipstr_synth hipstr_synth_t;
hipstr_synth_t* hipstr_synth_create(const hipstr_synth_cfg_t* cfg);
const hipstr_synth_view_t* hipstr_synth_view(const hipstr_synth_t* s);
void hipstr_synth_destroy(hipstr_synth_t* s);

#ifdef __cplusplus
}
#endif
#endif
