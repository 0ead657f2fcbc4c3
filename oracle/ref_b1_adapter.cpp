/*
 * ref_b1_adapter.cpp -- the seam-B1 binding of INTEGRATION.md, COMPILED against the reference's own headers.
 *
 * B200SeqStutterGenotyper has the constructor / genotype() / write_vcf_record() signatures of the reference's
 * SeqStutterGenotyper (src/seq_stutter_genotyper.h:143-146, 179-181, 189) plus the device context; a maintainer swaps
 * the type at src/genotyper_bam_processor.cpp:229 and nothing else in analyze_reads_and_phasing changes.  It takes the
 * reference's std::vector<Alignment>, RegionGroup, StutterModel* and per-sample log_p1 / log_p2 vectors, flattens them
 * into hipstr_locus_reads_t (include/hipstr_b200.h), runs the product through the C-ABI (hipstr_genotyper_*), and hands
 * the record to the reference's VCFWriter::add_vcf_record (src/vcf_writer.h:76).
 *
 * The extern "C" entry at the bottom is the test driver: it builds the reference objects from the flat test inputs
 * exactly like ref_genotyper_harness.cpp does for the reference's own class and runs the caller's sequence
 * (genotyper_bam_processor.cpp:229-246).  Built by oracle/Makefile into oracle/_ref/libhipstr_b1_adapter.so against
 * the reference objects; the hipstr_* symbols come from hipstr_b200/libhipstr_b200.so at load time.  No reference
 * source is modified or copied.
 */
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "vcf_writer.h"
#include "genotyper.h"
#include "seq_stutter_genotyper.h"
#include "mathops.h"
#include "region.h"
#include "vcf_reader.h"
#include "stutter_model.h"

#include "../include/hipstr_b200.h"

class B200SeqStutterGenotyper {
 public:
  /* SeqStutterGenotyper(region_group, haploid, reassemble_flanks, alignments, log_p1, log_p2, sample_names, chrom_seq,
   * stutter_models, ref_vcf, logger) -- seq_stutter_genotyper.h:143-146.  `alignments` is the flattened, sample-major
   * list matching log_p1[i][j] (genotyper.h:104-112); both mates of a pair are adjacent with equal names (.cpp:499). */
  B200SeqStutterGenotyper(hipstr_ctx_t* ctx, const RegionGroup& region_group, bool haploid, bool reassemble_flanks,
                          std::vector<Alignment>& alignments, std::vector< std::vector<double> >& log_p1,
                          std::vector< std::vector<double> >& log_p2, const std::vector<std::string>& sample_names,
                          const std::string& chrom_seq, std::vector<StutterModel*>& stutter_models, VCF::VCFReader* ref_vcf,
                          std::ostream& logger)
      : region_group_(region_group), reassemble_flanks_(reassemble_flanks), g_(NULL) {
    if (ref_vcf != NULL) { logger << "B200SeqStutterGenotyper: reference panels go through hipstr_genotyper_create_with_ref_alleles" << std::endl; return; }
    if (region_group.num_regions() != 1 || stutter_models.size() != 1) { logger << "B200SeqStutterGenotyper: one STR region per locus" << std::endl; return; }
    const Region& region = region_group.regions()[0];
    const int32_t R = (int32_t)alignments.size(), S = (int32_t)sample_names.size();
    std::vector<int32_t> locus_read_off(2, 0), locus_sample_off(2, 0), seq_off(1, 0), start(R), stop(R), cigar_off(1, 0), cigar_len, label(R), name_id(R);
    std::vector<double> p1(R), p2(R);
    std::vector<uint8_t> rev(R), use(R), hap(1, haploid ? 1 : 0);
    std::string bases, quals;
    std::vector<char> cigar_type;
    locus_read_off[1] = R;
    locus_sample_off[1] = S;
    int32_t r = 0;
    for (int32_t s = 0; s < S; s++)
      for (size_t j = 0; j < log_p1[s].size(); j++, r++) { label[r] = s; p1[r] = log_p1[s][j]; p2[r] = log_p2[s][j]; }
    if (r != R) { logger << "B200SeqStutterGenotyper: alignments and log_p1 disagree" << std::endl; return; }
    for (r = 0; r < R; r++) {
      const Alignment& a = alignments[r];
      bases += a.get_sequence();
      quals += a.get_base_qualities();
      seq_off.push_back((int32_t)bases.size());
      start[r] = a.get_start();
      stop[r] = a.get_stop();
      for (const CigarElement& c : a.get_cigar_list()) { cigar_type.push_back(c.get_type()); cigar_len.push_back(c.get_num()); }
      cigar_off.push_back((int32_t)cigar_type.size());
      name_id[r] = (r > 0 && a.get_name() == alignments[r - 1].get_name()) ? name_id[r - 1] : r;   // equal ids on adjacent reads = mates
      rev[r] = a.is_from_reverse_strand() ? 1 : 0;
      use[r] = a.use_for_hap_generation(0) ? 1 : 0;
    }
    hipstr_locus_reads_t rd;
    std::memset(&rd, 0, sizeof(rd));
    rd.locus_read_off = locus_read_off.data(); rd.locus_sample_off = locus_sample_off.data(); rd.read_seq_off = seq_off.data();
    rd.bases = bases.c_str(); rd.quals = quals.c_str(); rd.read_start = start.data(); rd.cigar_off = cigar_off.data();
    rd.cigar_type = cigar_type.data(); rd.cigar_len = cigar_len.data(); rd.sample_label = label.data(); rd.name_id = name_id.data();
    rd.log_p1 = p1.data(); rd.log_p2 = p2.data(); rd.haploid = hap.data(); rd.rev_strand = rev.data(); rd.read_stop = stop.data();
    rd.use_for_haps = use.data();
    const StutterModel* m = stutter_models[0];
    const double stutter[6] = {m->get_parameter(true, 'P'), m->get_parameter(true, 'U'), m->get_parameter(true, 'D'),
                               m->get_parameter(false, 'P'), m->get_parameter(false, 'U'), m->get_parameter(false, 'D')};
    const int32_t rs = region.start(), re = region.stop(), period = region.period();
    const char* chrom = chrom_seq.c_str();
    if (hipstr_genotyper_create_from_reads(ctx, 1, &rs, &re, &period, &chrom, stutter, &rd, &g_) != HIPSTR_OK) g_ = NULL;   // inputs are copied, like alns_ = alignments
  }
  ~B200SeqStutterGenotyper() { if (g_) hipstr_genotyper_destroy(g_); }

  /* bool genotype(max_total_haplotypes, max_flank_haplotypes, min_flank_freq, logger) -- seq_stutter_genotyper.h:189 */
  bool genotype(int max_total_haplotypes, int max_flank_haplotypes, double min_flank_freq, std::ostream& logger) {
    if (!g_) return false;
    uint8_t ok = 0;
    if (hipstr_genotyper_genotype(g_, max_total_haplotypes, max_flank_haplotypes, min_flank_freq, reassemble_flanks_ ? 1 : 0, &ok) != HIPSTR_OK) {
      logger << "hipstr_genotyper_genotype: " << hipstr_genotyper_last_error(g_) << std::endl;
      return false;
    }
    std::vector<char> log(1 << 16);
    const int32_t n = hipstr_genotyper_locus_log(g_, 0, log.data(), (int32_t)log.size());
    if (n > 0) logger << std::string(log.data(), (size_t)n);
    return ok != 0;
  }

  /* void write_vcf_record(sample_names, chrom_seq, output_viz, viz_left_alns, html_output, vcf_writer, logger) -- .h:179-181 */
  void write_vcf_record(const std::vector<std::string>& sample_names, const std::string& chrom_seq, bool output_viz, bool viz_left_alns,
                        std::ostream& html_output, VCFWriter* vcf_writer, std::ostream& logger) {
    (void)output_viz; (void)viz_left_alns; (void)html_output;   // the HTML visualisation is not produced
    const Region& region = region_group_.regions()[0];
    std::vector<const char*> names;
    for (const std::string& s : sample_names) names.push_back(s.c_str());
    const char* chrom = region.chrom().c_str();
    const char* name = region.name().c_str();
    const char* seq = chrom_seq.c_str();
    const int32_t rs = region.start(), re = region.stop(), period = region.period();
    hipstr_vcf_loci_t vl;
    std::memset(&vl, 0, sizeof(vl));
    vl.chrom = &chrom; vl.name = &name; vl.region_start = &rs; vl.region_stop = &re; vl.period = &period; vl.chrom_seq = &seq;
    vl.locus_sample_names = names.data(); vl.n_out_samples = (int32_t)names.size(); vl.out_sample_names = names.data();
    if (hipstr_genotyper_write_vcf(g_, &vl, NULL) != HIPSTR_OK) { logger << "hipstr_genotyper_write_vcf: " << hipstr_genotyper_last_error(g_) << std::endl; return; }
    int32_t pos = 0;
    const int32_t need = -hipstr_genotyper_locus_record(g_, 0, &pos, NULL, 0);
    std::vector<char> text((size_t)std::max(need, 1));
    if (hipstr_genotyper_locus_record(g_, 0, &pos, text.data(), (int32_t)text.size()) > 0)
      vcf_writer->add_vcf_record(region.chrom(), pos, std::string(text.data()));
  }

 private:
  RegionGroup region_group_;
  bool reassemble_flanks_;
  hipstr_genotyper_t* g_;
};

extern "C" {

/* The caller's sequence of genotyper_bam_processor.cpp:229-246 with the adapter in the reference class's place.  Flat
 * inputs as ref_sg_create (ref_genotyper_harness.cpp).  The record goes through the REFERENCE's VCFWriter into vcf_path
 * (BGZF, like the reference program's output).  Returns 1 if genotype() succeeded, 0 if not, < 0 on an error. */
int32_t b1_adapter_run(int32_t device, int32_t n_samples, int32_t n_reads, const int32_t* sample_label, const int32_t* name_id,
                       const int32_t* read_start, const int32_t* read_stop, const int32_t* seq_off, const char* bases, const char* quals,
                       const int32_t* cigar_off, const char* cigar_type, const int32_t* cigar_len, const double* log_p1, const double* log_p2,
                       const char* chrom_seq_c, int32_t region_start, int32_t region_stop, int32_t period, const double* stutter,
                       int32_t haploid, const uint8_t* rev_strand, const char* vcf_path) {
  precompute_integer_logs();
  hipstr_ctx_t* ctx = NULL;
  if (hipstr_create(device, &ctx) != HIPSTR_OK) return -1;
  std::vector<std::string> names;
  std::vector<Alignment> alns;
  std::vector<std::vector<double> > p1(n_samples), p2(n_samples);
  for (int s = 0; s < n_samples; s++) names.push_back("S" + std::to_string(s));
  for (int r = 0; r < n_reads; r++) {
    std::string seq(bases + seq_off[r], bases + seq_off[r + 1]), q(quals + seq_off[r], quals + seq_off[r + 1]);
    std::string gapped;
    int32_t pos = read_start[r];
    size_t k = 0;
    std::vector<CigarElement> cig;
    for (int c = cigar_off[r]; c < cigar_off[r + 1]; c++) {
      cig.push_back(CigarElement(cigar_type[c], cigar_len[c]));
      if (cigar_type[c] == 'D') { gapped.append(cigar_len[c], '-'); pos += cigar_len[c]; }
      else {
        gapped.append(seq, k, cigar_len[c]);
        k += cigar_len[c];
        if (cigar_type[c] != 'I') pos += cigar_len[c];
      }
    }
    Alignment a(read_start[r], read_stop ? read_stop[r] : pos - 1, rev_strand != NULL && rev_strand[r] != 0, "r" + std::to_string(name_id[r]), q, seq, gapped);
    a.set_cigar_list(cig);
    a.set_hap_gen_info(std::vector<bool>(1, true));
    alns.push_back(a);
    p1[sample_label[r]].push_back(log_p1[r]);
    p2[sample_label[r]].push_back(log_p2[r]);
  }
  const std::string chrom_seq(chrom_seq_c);
  Region region("chrS", region_start, region_stop, period, "STR");
  RegionGroup group(region);
  std::vector<StutterModel*> models(1, new StutterModel(stutter[0], stutter[1], stutter[2], stutter[3], stutter[4], stutter[5], period));
  std::ostringstream logger, html;
  VCFWriter writer;
  writer.open(vcf_path);
  int32_t result;
  {
    // genotyper_bam_processor.cpp:229-246, the reference's lines with the class name changed
    B200SeqStutterGenotyper* seq_genotyper = new B200SeqStutterGenotyper(ctx, group, haploid != 0, true, alns, p1, p2, names, chrom_seq, models, NULL, logger);
    if (seq_genotyper->genotype(1000, 4, 0.01, logger)) {
      seq_genotyper->write_vcf_record(names, chrom_seq, false, false, html, &writer, logger);
      result = 1;
    } else
      result = 0;
    delete seq_genotyper;
  }
  writer.close();
  delete models[0];
  hipstr_destroy(ctx);
  return result;
}

}  // extern "C"
