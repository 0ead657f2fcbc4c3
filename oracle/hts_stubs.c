/*
 * hts_stubs.c -- TEST INFRASTRUCTURE ONLY.
 * The reference's genotyper sources reference htslib's BAM / VCF / tabix entry points through
 * BamAlignment, VCF::VCFReader and bgzf streams.  The path the parity harness drives
 * (SeqStutterGenotyper with ref_vcf == NULL, records read from VCFWriter's heap) never calls
 * them, so htslib is not built; each symbol resolves to a stub that aborts if it is ever reached.
 * (bam_init1 / bam_destroy1 / bam_copy1 are real: see bam_min.c.)
 */
#include <stdlib.h>
#define HTS_STUB(name) void name(void) { abort(); }
HTS_STUB(bam_endpos) HTS_STUB(bam_hdr_destroy)
HTS_STUB(bcf_get_format_values) HTS_STUB(bcf_get_info) HTS_STUB(bcf_get_info_values) HTS_STUB(bcf_hdr_read)
HTS_STUB(bcf_unpack) HTS_STUB(cram_load_reference) HTS_STUB(hts_close) HTS_STUB(hts_get_bgzfp)
HTS_STUB(hts_idx_destroy) HTS_STUB(hts_itr_destroy) HTS_STUB(hts_itr_next) HTS_STUB(hts_itr_query)
HTS_STUB(hts_itr_querys) HTS_STUB(hts_open) HTS_STUB(sam_hdr_read) HTS_STUB(sam_index_load)
HTS_STUB(sam_itr_querys) HTS_STUB(tbx_index_load) HTS_STUB(tbx_name2id) HTS_STUB(tbx_readrec)
HTS_STUB(tbx_seqnames) HTS_STUB(vcf_parse) HTS_STUB(bgzf_open) HTS_STUB(bgzf_close) HTS_STUB(bgzf_getc)
HTS_STUB(bgzf_write)
