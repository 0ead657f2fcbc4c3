/* flank_assembler.cpp -- see flank_assembler.h. */
#include "flank_assembler.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <set>

namespace hipstr {

FlankAssembler::FlankAssembler(int k, const std::string& ref_seq) : k_(k), num_strings_(0) {
  source_kmer_ = ref_seq.substr(0, k);
  sink_kmer_ = ref_seq.substr(ref_seq.size() - k, k);
  add_string(ref_seq, 2);
  for (Edge& e : edges_) e.from_ref = true;
}

int FlankAssembler::node(std::string_view kmer) {
  auto it = node_of_.find(kmer);
  if (it != node_of_.end()) return it->second;
  const int id = (int)labels_.size();
  labels_.emplace_back(kmer);
  node_of_.emplace(std::string_view(labels_.back()), id);
  arriving_.emplace_back();
  departing_.emplace_back();
  return id;
}

void FlankAssembler::increment_edge(int s, int d, int delta) {
  for (int e : arriving_[d])
    if (edges_[e].source == s) { edges_[e].weight += delta; return; }
  const int id = (int)edges_.size();
  edges_.push_back(Edge{s, d, delta, false});
  departing_[s].push_back(id);
  arriving_[d].push_back(id);
}

void FlankAssembler::add_string(std::string_view seq, int weight, int copies) {
  if ((int)seq.size() <= k_) return;
  num_strings_ += copies;
  // consecutive edges share a node: every k-mer is looked up once (nodes are still created source first, in string order)
  int from = node(seq.substr(0, k_));
  for (size_t i = 1; i + k_ <= seq.size(); i++) {
    const int to = node(seq.substr(i, k_));
    increment_edge(from, to, weight * copies);
    from = to;
  }
}

void FlankAssembler::prune_edges(double min_edge_freq, int min_weight) {
  min_weight = std::max(min_weight, (int)std::ceil(min_edge_freq * num_strings_));
  const int n_nodes = (int)labels_.size(), n_edges = (int)edges_.size();
  std::vector<int> new_edge_id(n_edges, -1);
  std::vector<bool> keep_node(n_nodes, false);
  keep_node[node(source_kmer_)] = true;
  keep_node[node(sink_kmer_)] = true;
  std::vector<Edge> kept;
  for (int e = 0; e < n_edges; e++) {
    if (!edges_[e].from_ref && edges_[e].weight < min_weight) continue;
    new_edge_id[e] = (int)kept.size();
    kept.push_back(edges_[e]);
    keep_node[edges_[e].source] = true;
    keep_node[edges_[e].destination] = true;
  }
  // nodes that still touch an edge (plus source and sink) are renumbered in their old order
  std::vector<int> new_node_id(n_nodes, -1);
  std::deque<std::string> labels;
  std::vector<std::vector<int> > arriving, departing;
  node_of_.clear();
  for (int v = 0; v < n_nodes; v++) {
    if (!keep_node[v]) continue;
    new_node_id[v] = (int)labels.size();
    labels.push_back(labels_[v]);
    std::vector<int> in, out;
    for (int e : arriving_[v]) if (new_edge_id[e] >= 0) in.push_back(new_edge_id[e]);
    for (int e : departing_[v]) if (new_edge_id[e] >= 0) out.push_back(new_edge_id[e]);
    arriving.push_back(in);
    departing.push_back(out);
  }
  for (Edge& e : kept) { e.source = new_node_id[e.source]; e.destination = new_node_id[e.destination]; }
  edges_.swap(kept);
  labels_.swap(labels);
  for (size_t v = 0; v < labels_.size(); v++) node_of_.emplace(std::string_view(labels_[v]), (int)v);
  arriving_.swap(arriving);
  departing_.swap(departing);
}

bool FlankAssembler::has_cycles() const {   // Kahn's algorithm: a cycle leaves nodes unprocessed
  const int n = (int)labels_.size();
  std::vector<int> pending(n);
  std::vector<int> ready;
  int left = 0;
  for (int v = 0; v < n; v++) {
    pending[v] = (int)arriving_[v].size();
    if (pending[v] == 0) ready.push_back(v);
    else left++;
  }
  while (!ready.empty()) {
    const int v = ready.back();
    ready.pop_back();
    for (int e : departing_[v]) {
      const int c = edges_[e].destination;
      if (--pending[c] == 0) { ready.push_back(c); left--; }
    }
  }
  return left != 0;
}

bool FlankAssembler::is_source_ok() {
  const int s = node(source_kmer_);
  return !departing_[s].empty() && arriving_[s].empty();
}
bool FlankAssembler::is_sink_ok() {
  const int s = node(sink_kmer_);
  return !arriving_[s].empty() && departing_[s].empty();
}

/* k-mers one substitution away from `kmer` that are present and are themselves sources / sinks */
void FlankAssembler::alt_kmer_nodes(std::string kmer, bool source, bool sink, std::vector<int>& nodes) {
  const char bases[4] = {'A', 'C', 'G', 'T'};
  for (size_t i = 0; i < kmer.size(); i++) {
    const char orig = kmer[i];
    for (char b : bases) {
      if (b == orig) continue;
      kmer[i] = b;
      auto it = node_of_.find(kmer);
      if (it == node_of_.end()) continue;
      if (source && !arriving_[it->second].empty()) continue;
      if (sink && !departing_[it->second].empty()) continue;
      nodes.push_back(it->second);
    }
    kmer[i] = orig;
  }
}

void FlankAssembler::enumerate_paths(int min_weight, int max_paths, std::vector<std::pair<std::string, int> >& paths) {
  struct Path { int parent, node, min_weight; };
  std::vector<Path> all;
  std::vector<int> heap;   // indices into `all`; the path with the largest bottleneck weight is popped first
  auto lighter = [&all](int a, int b) { return all[a].min_weight < all[b].min_weight; };
  const int source = node(source_kmer_), sink = node(sink_kmer_);
  all.push_back(Path{-1, source, 1000000});
  heap.push_back(0);
  std::make_heap(heap.begin(), heap.end(), lighter);
  std::vector<int> alt;
  alt_kmer_nodes(source_kmer_, true, false, alt);
  for (int v : alt) {
    all.push_back(Path{-1, v, 1000000});
    heap.push_back((int)all.size() - 1);
    std::push_heap(heap.begin(), heap.end(), lighter);
  }
  std::set<int> sinks;
  sinks.insert(sink);
  alt.clear();
  alt_kmer_nodes(sink_kmer_, false, true, alt);
  sinks.insert(alt.begin(), alt.end());
  while (!heap.empty()) {
    if ((int)paths.size() == max_paths) break;
    std::pop_heap(heap.begin(), heap.end(), lighter);
    const int best = heap.back();
    heap.pop_back();
    if (sinks.count(all[best].node)) {
      std::string seq;   // first base of every k-mer on the path, then the rest of the last k-mer
      std::vector<int> chain;
      for (int p = best; p != -1; p = all[p].parent) chain.push_back(p);
      for (size_t i = chain.size(); i-- > 1;) seq += labels_[all[chain[i]].node][0];
      seq += labels_[all[best].node];
      paths.emplace_back(seq, all[best].min_weight);
    }
    const std::vector<int> out = departing_[all[best].node];
    for (int e : out) {
      if (edges_[e].weight < min_weight) continue;
      all.push_back(Path{best, edges_[e].destination, std::min(all[best].min_weight, edges_[e].weight)});
      heap.push_back((int)all.size() - 1);
      std::push_heap(heap.begin(), heap.end(), lighter);
    }
  }
}

bool FlankAssembler::calc_kmer_length(const std::string& ref_seq, int min_kmer, int max_kmer, int& kmer) {
  for (kmer = min_kmer; kmer <= max_kmer; kmer++) {
    FlankAssembler graph(kmer, ref_seq);
    if (!graph.has_cycles()) return true;
  }
  return false;
}

}  // namespace hipstr

/* Exported for checking on its own (include/hipstr_b200.h): the per-sample assembly step of assemble_flanks
 * (seq_stutter_genotyper.cpp:76-97). */
extern "C" int32_t hipstr_flank_assemble(const char* ref_seq, int32_t n_seqs, const char* const* seqs, int32_t min_kmer, int32_t max_kmer,
                                         int32_t* k_used, int32_t max_paths, int32_t path_cap, char* paths, int32_t* weights) {
  if (!ref_seq || (n_seqs > 0 && !seqs) || !k_used || !paths || !weights) return -2;
  const std::string ref(ref_seq);
  const int max_k = std::min(max_kmer, ref.empty() ? -1 : (int)ref.size() - 1);
  int kmer_length;
  if (!hipstr::FlankAssembler::calc_kmer_length(ref, min_kmer, max_k, kmer_length)) return -1;   // flank too repetitive
  for (int k = kmer_length; k <= max_k; k++) {
    hipstr::FlankAssembler assembler(k, ref);
    for (int i = 0; i < n_seqs; i++) {
      const std::string s(seqs[i]);
      if (!s.empty()) assembler.add_string(s);
    }
    assembler.prune_edges(0.02, 2);
    if (!assembler.has_cycles() && assembler.is_source_ok() && assembler.is_sink_ok()) {
      std::vector<std::pair<std::string, int> > found;
      assembler.enumerate_paths(2, max_paths, found);
      *k_used = k;
      for (size_t i = 0; i < found.size(); i++) {
        if ((int32_t)found[i].first.size() + 1 > path_cap) return -2;
        std::memcpy(paths + i * (size_t)path_cap, found[i].first.c_str(), found[i].first.size() + 1);
        weights[i] = found[i].second;
      }
      return (int32_t)found.size();
    }
  }
  return -3;   // cyclic for every k: the sample is marked FLANK_ASSEMBLY_CYCLIC
}
