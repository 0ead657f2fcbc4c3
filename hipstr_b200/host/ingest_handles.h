/* ingest_handles.h -- the opaque handles of the ingestion C-ABI (ingest_capi.cpp), shared with the region driver. */
#ifndef HIPSTR_B200_INGEST_HANDLES_H_
#define HIPSTR_B200_INGEST_HANDLES_H_

#include <memory>
#include <string>
#include <vector>

#include "bam_reader.h"
#include "read_filter.h"

using hipstr::BamRecord;

struct hipstr_bam_reader {
  std::vector<std::unique_ptr<hipstr::BamFile> > files;
  std::vector<std::string> paths;
  std::string error;
};
struct hipstr_bam_records {
  std::vector<BamRecord> records;
  std::vector<std::string> ref_names, file_names;
};
struct hipstr_filtered_reads {
  hipstr::FilteredReads reads;
  std::string adapter_stats;
  // flat view (hipstr_filtered_reads_view)
  std::vector<int32_t> sample_entry_off, entry_aln_off, entry_snp_set, aln_pos, aln_end, aln_seq_off, aln_cigar_off, cigar_len, aln_flag,
      entry_name_off;
  std::string bases, quals, cigar_type, passes, entry_names;
  std::vector<const char*> sample_names;
};

#endif
