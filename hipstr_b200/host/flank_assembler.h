/*
 * flank_assembler.h -- local assembly of the sequence flanking an STR from the reads of ONE sample:
 * host-side restatement of the reference's DebruijnGraph / DirectedGraph / DebruijnPath
 * (src/debruijn_graph.{h,cpp}, src/directed_graph.{h,cpp}) as used by
 * SeqStutterGenotyper::assemble_flanks (src/seq_stutter_genotyper.cpp:40-217).
 *
 * Index-based: k-mers are numbered in order of first appearance, edges live in one array with
 * per-node in/out adjacency lists that keep insertion order -- the order is part of the result
 * (path enumeration pops a heap whose ties are broken by insertion order), so it is preserved
 * through pruning exactly as the reference's pointer-based graph does.
 */
#ifndef HIPSTR_B200_FLANK_ASSEMBLER_H_
#define HIPSTR_B200_FLANK_ASSEMBLER_H_

#include <deque>
#include <string>
#include <string_view>
#include <unordered_map>
#include <utility>
#include <vector>

namespace hipstr {

class FlankAssembler {
 public:
  /* DebruijnGraph(k, ref_seq): the reference path enters with weight 2 and its edges are never pruned. */
  FlankAssembler(int k, const std::string& ref_seq);
  /* debruijn_graph.cpp:31-45; `copies` identical strings at once (same graph as adding them one by one) */
  void add_string(std::string_view seq, int weight = 1, int copies = 1);
  void prune_edges(double min_edge_freq, int min_weight);         /* :47-60, 62-121 */
  bool has_cycles() const;                                        /* directed_graph.cpp:29-64 */
  bool is_source_ok();                                            /* debruijn_graph.cpp:12-15 */
  bool is_sink_ok();                                              /* :17-20 */
  /* Best-first enumeration of source->sink paths by bottleneck weight (:151-199). */
  void enumerate_paths(int min_weight, int max_paths, std::vector<std::pair<std::string, int> >& paths);
  /* DebruijnGraph::calc_kmer_length (:22-29) */
  static bool calc_kmer_length(const std::string& ref_seq, int min_kmer, int max_kmer, int& kmer);

 private:
  struct Edge { int source, destination, weight; bool from_ref; };
  int k_;
  std::string source_kmer_, sink_kmer_;
  int num_strings_;
  std::deque<std::string> labels_;                /* node id -> k-mer (stable storage: node_of_ keys view into it) */
  std::unordered_map<std::string_view, int> node_of_;
  std::vector<Edge> edges_;
  std::vector<std::vector<int> > arriving_, departing_;   /* node id -> edge ids, insertion order */

  int node(std::string_view kmer);                /* get_node: creates the node when absent */
  void increment_edge(int from_node, int to_node, int delta);
  void alt_kmer_nodes(std::string kmer, bool source, bool sink, std::vector<int>& nodes);
};

}  // namespace hipstr
#endif
