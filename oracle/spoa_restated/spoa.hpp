// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked into the product library.
//
// Restatement of the published algorithm of rvaser/spoa tag 1.1.3 (MIT), the third-party
// dependency HASLR links for its POA consensus (reference: src/haslr_assemble/Makefile:1,41-47;
// call sites src/haslr_assemble/src/Assemble.cpp:499,500,539,540,554). The SPOA sources are NOT in
// /root/reference and there is no network, so this file is written from the published algorithm
// (global Needleman-Wunsch of a sequence against a partial-order graph, linear gaps; graph update;
// DFS topological sort with aligned nodes kept adjacent; heaviest-bundle consensus with branch
// completion).
//
// *** PARITY UNPINNED ***: the reference holds no golden vectors for this boundary
// and real SPOA cannot be run here. What this header guarantees is the API surface the reference
// calls, and the SPOA 1.1.3 semantics as restated in SURVEY.md §8(c).
//
// Only what HASLR uses is provided: AlignmentType::kNW with linear gaps (match 5, mismatch -4,
// gap -8 at the call site), Graph::add_alignment with unit weights, Graph::generate_consensus.
#ifndef SPOA_RESTATED_HPP
#define SPOA_RESTATED_HPP

#include <cstdint>
#include <cstdlib>
#include <algorithm>
#include <limits>
#include <memory>
#include <stack>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#if defined(__SSE4_1__)
#include <smmintrin.h>
#endif

namespace spoa {

enum class AlignmentType { kSW, kNW, kOV };
using Alignment = std::vector<std::pair<std::int32_t, std::int32_t>>;

struct Edge {
    std::uint32_t begin_node_id;
    std::uint32_t end_node_id;
    std::int64_t total_weight;
};

struct Node {
    std::uint32_t id;
    std::uint32_t code;
    std::vector<std::uint32_t> in_edges;   // indices into Graph::edges_, insertion order
    std::vector<std::uint32_t> out_edges;  // indices into Graph::edges_, insertion order
    std::vector<std::uint32_t> aligned_nodes_ids;
};

class Graph {
public:
    Graph() : num_sequences_(0), num_codes_(0), coder_(256, -1), decoder_(256, -1) {}

    const std::vector<Node>& nodes() const { return nodes_; }
    const std::vector<Edge>& edges() const { return edges_; }
    const std::vector<std::uint32_t>& rank_to_node_id() const { return rank_to_node_id_; }
    std::uint32_t num_codes() const { return num_codes_; }
    std::int32_t coder(char c) const { return coder_[static_cast<unsigned char>(c)]; }
    char decoder(std::uint32_t code) const { return static_cast<char>(decoder_[code]); }

    void add_alignment(const Alignment& alignment, const std::string& sequence, std::uint32_t weight = 1) {
        std::vector<std::uint32_t> weights(sequence.size(), weight);
        add_alignment(alignment, sequence.c_str(), static_cast<std::uint32_t>(sequence.size()), weights);
    }

    void add_alignment(const Alignment& alignment, const char* sequence, std::uint32_t sequence_size,
                       const std::vector<std::uint32_t>& weights) {
        if (sequence_size == 0) return;
        if (sequence_size != weights.size()) throw std::invalid_argument("[spoa restated] sequence/weights size mismatch");
        for (std::uint32_t i = 0; i < sequence_size; ++i) {
            unsigned char c = static_cast<unsigned char>(sequence[i]);
            if (coder_[c] == -1) {
                coder_[c] = static_cast<std::int32_t>(num_codes_);
                decoder_[num_codes_] = c;
                ++num_codes_;
            }
        }
        if (alignment.empty()) {  // first sequence (or unalignable): a fresh chain
            std::int32_t begin_node_id = add_sequence(sequence, weights, 0, sequence_size);
            ++num_sequences_;
            sequences_begin_nodes_ids_.push_back(static_cast<std::uint32_t>(begin_node_id));
            topological_sort();
            return;
        }
        std::vector<std::uint32_t> valid_seq_ids;
        for (const auto& it : alignment)
            if (it.second != -1) valid_seq_ids.push_back(static_cast<std::uint32_t>(it.second));

        std::uint32_t tmp = static_cast<std::uint32_t>(nodes_.size());
        std::int32_t begin_node_id = add_sequence(sequence, weights, 0, valid_seq_ids.front());
        std::int32_t head_node_id = tmp == nodes_.size() ? -1 : static_cast<std::int32_t>(nodes_.size()) - 1;
        std::int32_t tail_node_id = add_sequence(sequence, weights, valid_seq_ids.back() + 1, sequence_size);

        std::int32_t new_node_id = -1;
        float prev_weight = head_node_id == -1 ? 0 : static_cast<float>(weights[valid_seq_ids.front() - 1]);

        for (std::uint32_t i = 0; i < alignment.size(); ++i) {
            if (alignment[i].second == -1) continue;
            char letter = sequence[alignment[i].second];
            std::uint32_t code = static_cast<std::uint32_t>(coder_[static_cast<unsigned char>(letter)]);
            if (alignment[i].first == -1) {
                new_node_id = static_cast<std::int32_t>(add_node(code));
            } else {
                std::uint32_t nid = static_cast<std::uint32_t>(alignment[i].first);
                if (nodes_[nid].code == code) {
                    new_node_id = alignment[i].first;
                } else {
                    std::int32_t aligned_to_node_id = -1;
                    for (std::uint32_t aid : nodes_[nid].aligned_nodes_ids) {
                        if (nodes_[aid].code == code) { aligned_to_node_id = static_cast<std::int32_t>(aid); break; }
                    }
                    if (aligned_to_node_id == -1) {
                        new_node_id = static_cast<std::int32_t>(add_node(code));
                        // copy: add_node may not reallocate inner vectors, but nodes_ may move
                        std::vector<std::uint32_t> others = nodes_[nid].aligned_nodes_ids;
                        for (std::uint32_t aid : others) {
                            nodes_[new_node_id].aligned_nodes_ids.push_back(aid);
                            nodes_[aid].aligned_nodes_ids.push_back(static_cast<std::uint32_t>(new_node_id));
                        }
                        nodes_[new_node_id].aligned_nodes_ids.push_back(nid);
                        nodes_[nid].aligned_nodes_ids.push_back(static_cast<std::uint32_t>(new_node_id));
                    } else {
                        new_node_id = aligned_to_node_id;
                    }
                }
            }
            if (begin_node_id == -1) begin_node_id = new_node_id;
            if (head_node_id != -1) {
                // both nodes contribute to the edge weight
                add_edge(static_cast<std::uint32_t>(head_node_id), static_cast<std::uint32_t>(new_node_id),
                         static_cast<std::int64_t>(prev_weight + weights[alignment[i].second]));
            }
            head_node_id = new_node_id;
            prev_weight = static_cast<float>(weights[alignment[i].second]);
        }
        if (tail_node_id != -1) {
            add_edge(static_cast<std::uint32_t>(head_node_id), static_cast<std::uint32_t>(tail_node_id),
                     static_cast<std::int64_t>(prev_weight + weights[valid_seq_ids.back() + 1]));
        }
        ++num_sequences_;
        sequences_begin_nodes_ids_.push_back(static_cast<std::uint32_t>(begin_node_id));
        topological_sort();
    }

    std::string generate_consensus() {
        traverse_heaviest_bundle();
        std::string consensus_str;
        for (std::uint32_t id : consensus_) consensus_str += decoder(nodes_[id].code);
        return consensus_str;
    }

private:
    std::uint32_t add_node(std::uint32_t code) {
        std::uint32_t node_id = static_cast<std::uint32_t>(nodes_.size());
        nodes_.push_back(Node{node_id, code, {}, {}, {}});
        return node_id;
    }

    void add_edge(std::uint32_t begin_node_id, std::uint32_t end_node_id, std::int64_t weight) {
        for (std::uint32_t e : nodes_[begin_node_id].out_edges) {
            if (edges_[e].end_node_id == end_node_id) { edges_[e].total_weight += weight; return; }
        }
        std::uint32_t e = static_cast<std::uint32_t>(edges_.size());
        edges_.push_back(Edge{begin_node_id, end_node_id, weight});
        nodes_[begin_node_id].out_edges.push_back(e);
        nodes_[end_node_id].in_edges.push_back(e);
    }

    std::int32_t add_sequence(const char* sequence, const std::vector<std::uint32_t>& weights,
                              std::uint32_t begin, std::uint32_t end) {
        if (begin == end) return -1;
        std::int32_t first_node_id = static_cast<std::int32_t>(
            add_node(static_cast<std::uint32_t>(coder_[static_cast<unsigned char>(sequence[begin])])));
        std::uint32_t node_id;
        for (std::uint32_t i = begin + 1; i < end; ++i) {
            node_id = add_node(static_cast<std::uint32_t>(coder_[static_cast<unsigned char>(sequence[i])]));
            // both nodes contribute to edge weight
            add_edge(node_id - 1, node_id, static_cast<std::int64_t>(weights[i - 1]) + weights[i]);
        }
        return first_node_id;
    }

    void topological_sort() {
        rank_to_node_id_.clear();
        // 0 - unmarked, 1 - temporarily marked, 2 - permanently marked
        std::vector<std::uint8_t> node_marks(nodes_.size(), 0);
        std::vector<bool> check_aligned_nodes(nodes_.size(), true);
        std::stack<std::uint32_t> nodes_to_visit;
        for (std::uint32_t i = 0; i < nodes_.size(); ++i) {
            if (node_marks[i] != 0) continue;
            nodes_to_visit.push(i);
            while (!nodes_to_visit.empty()) {
                std::uint32_t node_id = nodes_to_visit.top();
                bool valid = true;
                if (node_marks[node_id] != 2) {
                    for (std::uint32_t e : nodes_[node_id].in_edges) {
                        std::uint32_t b = edges_[e].begin_node_id;
                        if (node_marks[b] != 2) { nodes_to_visit.push(b); valid = false; }
                    }
                    if (check_aligned_nodes[node_id]) {
                        for (std::uint32_t aid : nodes_[node_id].aligned_nodes_ids) {
                            if (node_marks[aid] != 2) {
                                nodes_to_visit.push(aid);
                                check_aligned_nodes[aid] = false;
                                valid = false;
                            }
                        }
                    }
                    if (!(valid || node_marks[node_id] != 1)) throw std::logic_error("[spoa restated] graph is not a DAG");
                    if (valid) {
                        node_marks[node_id] = 2;
                        if (check_aligned_nodes[node_id]) {
                            rank_to_node_id_.push_back(node_id);
                            for (std::uint32_t aid : nodes_[node_id].aligned_nodes_ids) rank_to_node_id_.push_back(aid);
                        }
                    } else {
                        node_marks[node_id] = 1;
                    }
                }
                if (valid) nodes_to_visit.pop();
            }
        }
    }

    void traverse_heaviest_bundle() {
        std::vector<std::int32_t> predecessors(nodes_.size(), -1);
        std::vector<std::int64_t> scores(nodes_.size(), -1);
        std::uint32_t max_score_id = 0;
        for (std::uint32_t node_id : rank_to_node_id_) {
            for (std::uint32_t e : nodes_[node_id].in_edges) {
                const Edge& edge = edges_[e];
                if (scores[node_id] < edge.total_weight ||
                    (scores[node_id] == edge.total_weight &&
                     scores[predecessors[node_id]] <= scores[edge.begin_node_id])) {
                    scores[node_id] = edge.total_weight;
                    predecessors[node_id] = static_cast<std::int32_t>(edge.begin_node_id);
                }
            }
            if (predecessors[node_id] != -1) scores[node_id] += scores[predecessors[node_id]];
            if (scores[max_score_id] < scores[node_id]) max_score_id = node_id;
        }
        if (!nodes_[max_score_id].out_edges.empty()) {
            std::vector<std::uint32_t> node_id_to_rank(nodes_.size(), 0);
            for (std::uint32_t i = 0; i < nodes_.size(); ++i) node_id_to_rank[rank_to_node_id_[i]] = i;
            while (!nodes_[max_score_id].out_edges.empty())
                max_score_id = branch_completion(scores, predecessors, node_id_to_rank[max_score_id]);
        }
        consensus_.clear();
        while (predecessors[max_score_id] != -1) {
            consensus_.push_back(max_score_id);
            max_score_id = static_cast<std::uint32_t>(predecessors[max_score_id]);
        }
        consensus_.push_back(max_score_id);
        std::reverse(consensus_.begin(), consensus_.end());
    }

    std::uint32_t branch_completion(std::vector<std::int64_t>& scores, std::vector<std::int32_t>& predecessors,
                                    std::uint32_t rank) {
        std::uint32_t node_id = rank_to_node_id_[rank];
        for (std::uint32_t e : nodes_[node_id].out_edges) {
            for (std::uint32_t oe : nodes_[edges_[e].end_node_id].in_edges) {
                if (edges_[oe].begin_node_id != node_id) scores[edges_[oe].begin_node_id] = -1;
            }
        }
        std::int64_t max_score = 0;
        std::uint32_t max_score_id = 0;
        for (std::uint32_t i = rank + 1; i < rank_to_node_id_.size(); ++i) {
            std::uint32_t nid = rank_to_node_id_[i];
            scores[nid] = -1;
            predecessors[nid] = -1;
            for (std::uint32_t e : nodes_[nid].in_edges) {
                const Edge& edge = edges_[e];
                if (scores[edge.begin_node_id] == -1) continue;
                if (scores[nid] < edge.total_weight ||
                    (scores[nid] == edge.total_weight &&
                     scores[predecessors[nid]] <= scores[edge.begin_node_id])) {
                    scores[nid] = edge.total_weight;
                    predecessors[nid] = static_cast<std::int32_t>(edge.begin_node_id);
                }
            }
            if (predecessors[nid] != -1) scores[nid] += scores[predecessors[nid]];
            if (max_score < scores[nid]) { max_score = scores[nid]; max_score_id = nid; }
        }
        return max_score_id;
    }

    std::uint32_t num_sequences_;
    std::uint32_t num_codes_;
    std::vector<std::int32_t> coder_;
    std::vector<std::int32_t> decoder_;
    std::vector<Node> nodes_;
    std::vector<Edge> edges_;
    std::vector<std::uint32_t> rank_to_node_id_;
    std::vector<std::uint32_t> sequences_begin_nodes_ids_;
    std::vector<std::uint32_t> consensus_;
};

// Global NW of a sequence against the graph, linear gap. Rows = nodes in topological order (row 0 is
// a virtual source), columns = sequence positions. Two fills producing identical matrices:
//  - scalar int32 (authoritative restatement),
//  - SSE4.1 int16 row-vectorised fill (8 lanes), chosen like SPOA's SIMD engine when the worst-case
//    |score| fits int16; used so that CPU-baseline timings resemble the SSE4.1 build the reference ships.
class AlignmentEngine {
public:
    AlignmentEngine(AlignmentType type, std::int8_t m, std::int8_t n, std::int8_t g)
        : type_(type), m_(m), n_(n), g_(g), use_simd_(true), cells_(0) {
        if (type != AlignmentType::kNW) throw std::invalid_argument("[spoa restated] only kNW is restated");
    }
    void set_simd(bool on) { use_simd_ = on; }
    std::uint64_t cells() const { return cells_; }  // DP cells filled so far (for GCUPS accounting)
    bool last_was_int16() const { return last_int16_; }

    Alignment align_sequence_with_graph(const std::string& sequence, const std::unique_ptr<Graph>& graph) {
        return align_sequence_with_graph(sequence.c_str(), static_cast<std::uint32_t>(sequence.size()), graph);
    }

    Alignment align_sequence_with_graph(const char* sequence, std::uint32_t sequence_size,
                                        const std::unique_ptr<Graph>& graph) {
        if (graph->nodes().empty() || sequence_size == 0) return Alignment();
        const auto& nodes = graph->nodes();
        const auto& edges = graph->edges();
        const auto& rank_to_node_id = graph->rank_to_node_id();
        const std::uint32_t H_rows = static_cast<std::uint32_t>(nodes.size()) + 1;
        const std::uint32_t W = sequence_size + 1;
        cells_ += static_cast<std::uint64_t>(H_rows) * W;

        node_id_to_rank_.assign(nodes.size(), 0);
        for (std::uint32_t i = 0; i < nodes.size(); ++i) node_id_to_rank_[rank_to_node_id[i]] = i;

        // sequence profile: profile[code * W + j] = score of aligning a node with `code` to sequence[j-1]
        const std::uint32_t ncodes = graph->num_codes();
        std::vector<std::int32_t> seq_codes(sequence_size);
        for (std::uint32_t j = 0; j < sequence_size; ++j) seq_codes[j] = graph->coder(sequence[j]);

        const std::int64_t max_penalty = std::max(std::max(std::abs((int)m_), std::abs((int)n_)), std::abs((int)g_));
        const std::int64_t longest_path = static_cast<std::int64_t>(nodes.size()) + 1 + sequence_size + 8;
        last_int16_ = false;
#if defined(__SSE4_1__)
        if (use_simd_ && max_penalty * longest_path < std::numeric_limits<std::int16_t>::max()) {
            last_int16_ = true;
            return align_int16(sequence_size, seq_codes, ncodes, nodes, edges, rank_to_node_id);
        }
#endif
        (void)max_penalty; (void)longest_path;
        return align_int32(sequence_size, seq_codes, ncodes, nodes, edges, rank_to_node_id);
    }

private:
    template <typename T>
    Alignment backtrack(const T* H, std::uint32_t W, std::uint32_t max_i, std::uint32_t max_j,
                        const std::vector<std::int32_t>& seq_codes, const std::vector<Node>& nodes,
                        const std::vector<Edge>& edges, const std::vector<std::uint32_t>& rank_to_node_id) {
        Alignment alignment;
        std::uint32_t i = max_i, j = max_j;
        std::uint32_t prev_i = 0, prev_j = 0;
        while (!(i == 0 && j == 0)) {
            const std::int32_t H_ij = H[static_cast<std::uint64_t>(i) * W + j];
            bool predecessor_found = false;
            if (i != 0 && j != 0) {
                const Node& node = nodes[rank_to_node_id[i - 1]];
                const std::int32_t match_cost = (static_cast<std::int32_t>(node.code) == seq_codes[j - 1]) ? m_ : n_;
                if (node.in_edges.empty()) {
                    if (H_ij == static_cast<std::int32_t>(H[j - 1]) + match_cost) { prev_i = 0; prev_j = j - 1; predecessor_found = true; }
                } else {
                    for (std::uint32_t e : node.in_edges) {
                        std::uint32_t pred_i = node_id_to_rank_[edges[e].begin_node_id] + 1;
                        if (H_ij == static_cast<std::int32_t>(H[static_cast<std::uint64_t>(pred_i) * W + j - 1]) + match_cost) {
                            prev_i = pred_i; prev_j = j - 1; predecessor_found = true; break;
                        }
                    }
                }
            }
            if (!predecessor_found && i != 0) {
                const Node& node = nodes[rank_to_node_id[i - 1]];
                if (node.in_edges.empty()) {
                    if (H_ij == static_cast<std::int32_t>(H[j]) + g_) { prev_i = 0; prev_j = j; predecessor_found = true; }
                } else {
                    for (std::uint32_t e : node.in_edges) {
                        std::uint32_t pred_i = node_id_to_rank_[edges[e].begin_node_id] + 1;
                        if (H_ij == static_cast<std::int32_t>(H[static_cast<std::uint64_t>(pred_i) * W + j]) + g_) {
                            prev_i = pred_i; prev_j = j; predecessor_found = true; break;
                        }
                    }
                }
            }
            if (!predecessor_found && j != 0 &&
                H_ij == static_cast<std::int32_t>(H[static_cast<std::uint64_t>(i) * W + j - 1]) + g_) {
                prev_i = i; prev_j = j - 1; predecessor_found = true;
            }
            if (!predecessor_found) throw std::logic_error("[spoa restated] backtrack found no predecessor");
            alignment.emplace_back(i == prev_i ? -1 : static_cast<std::int32_t>(rank_to_node_id[i - 1]),
                                   j == prev_j ? -1 : static_cast<std::int32_t>(j - 1));
            i = prev_i; j = prev_j;
        }
        std::reverse(alignment.begin(), alignment.end());
        return alignment;
    }

    Alignment align_int32(std::uint32_t L, const std::vector<std::int32_t>& seq_codes, std::uint32_t ncodes,
                          const std::vector<Node>& nodes, const std::vector<Edge>& edges,
                          const std::vector<std::uint32_t>& rank_to_node_id) {
        const std::uint32_t W = L + 1;
        const std::uint64_t rows = nodes.size() + 1;
        H32_.resize(rows * W);
        std::int32_t* H = H32_.data();
        std::vector<std::int32_t> profile(static_cast<std::uint64_t>(ncodes) * W);
        for (std::uint32_t c = 0; c < ncodes; ++c) {
            profile[static_cast<std::uint64_t>(c) * W] = 0;
            for (std::uint32_t j = 1; j < W; ++j)
                profile[static_cast<std::uint64_t>(c) * W + j] = (static_cast<std::int32_t>(c) == seq_codes[j - 1]) ? m_ : n_;
        }
        for (std::uint32_t j = 0; j < W; ++j) H[j] = static_cast<std::int32_t>(j) * g_;
        std::int32_t max_score = std::numeric_limits<std::int32_t>::min();
        std::uint32_t max_i = 0, max_j = 0;
        for (std::uint32_t r = 0; r < nodes.size(); ++r) {
            const Node& node = nodes[rank_to_node_id[r]];
            const std::int32_t* prof = &profile[static_cast<std::uint64_t>(node.code) * W];
            std::int32_t* Hr = H + static_cast<std::uint64_t>(r + 1) * W;
            // first column
            if (node.in_edges.empty()) {
                Hr[0] = g_;
            } else {
                std::int32_t best = std::numeric_limits<std::int32_t>::min();
                for (std::uint32_t e : node.in_edges)
                    best = std::max(best, H[static_cast<std::uint64_t>(node_id_to_rank_[edges[e].begin_node_id] + 1) * W]);
                Hr[0] = best + g_;
            }
            std::uint32_t pred_i = node.in_edges.empty() ? 0 : node_id_to_rank_[edges[node.in_edges[0]].begin_node_id] + 1;
            const std::int32_t* Hp = H + static_cast<std::uint64_t>(pred_i) * W;
            for (std::uint32_t j = 1; j < W; ++j) Hr[j] = std::max(Hp[j - 1] + prof[j], Hp[j] + g_);
            for (std::uint32_t p = 1; p < node.in_edges.size(); ++p) {
                pred_i = node_id_to_rank_[edges[node.in_edges[p]].begin_node_id] + 1;
                Hp = H + static_cast<std::uint64_t>(pred_i) * W;
                for (std::uint32_t j = 1; j < W; ++j) Hr[j] = std::max(Hp[j - 1] + prof[j], std::max(Hr[j], Hp[j] + g_));
            }
            for (std::uint32_t j = 1; j < W; ++j) Hr[j] = std::max(Hr[j - 1] + g_, Hr[j]);
            if (node.out_edges.empty() && max_score < Hr[W - 1]) { max_score = Hr[W - 1]; max_i = r + 1; max_j = W - 1; }
        }
        return backtrack<std::int32_t>(H, W, max_i, max_j, seq_codes, nodes, edges, rank_to_node_id);
    }

#if defined(__SSE4_1__)
    Alignment align_int16(std::uint32_t L, const std::vector<std::int32_t>& seq_codes, std::uint32_t ncodes,
                          const std::vector<Node>& nodes, const std::vector<Edge>& edges,
                          const std::vector<std::uint32_t>& rank_to_node_id) {
        const std::uint32_t W = L + 1;
        // padded pitch: 8 lanes of slack on the left (so the diagonal load at j-1 of the first vector
        // reads initialised memory) and up to 8 on the right.
        const std::uint32_t P = ((W + 7) / 8) * 8 + 8;
        const std::uint64_t rows = nodes.size() + 1;
        H16_.resize(rows * P + 8);
        std::int16_t* H = H16_.data() + 8;  // H[row*P + j], j in [-8, P-8)
        const std::int16_t kNegInf = std::numeric_limits<std::int16_t>::min() + 1024;
        std::vector<std::int16_t> profile(static_cast<std::uint64_t>(ncodes) * P, 0);
        for (std::uint32_t c = 0; c < ncodes; ++c)
            for (std::uint32_t j = 1; j < W; ++j)
                profile[static_cast<std::uint64_t>(c) * P + j] = (static_cast<std::int32_t>(c) == seq_codes[j - 1]) ? m_ : n_;
        for (std::uint32_t j = 0; j < P - 8; ++j) H[j] = j < W ? static_cast<std::int16_t>(static_cast<std::int32_t>(j) * g_) : kNegInf;
        const __m128i vg = _mm_set1_epi16(g_);
        const __m128i vg2 = _mm_set1_epi16(static_cast<std::int16_t>(2 * g_));
        const __m128i vg4 = _mm_set1_epi16(static_cast<std::int16_t>(4 * g_));
        const __m128i vneg = _mm_set1_epi16(kNegInf);
        const __m128i vramp = _mm_mullo_epi16(_mm_set_epi16(8, 7, 6, 5, 4, 3, 2, 1), vg);  // (k+1)*g
        const __m128i m1 = _mm_set_epi16(0, 0, 0, 0, 0, 0, 0, -1);
        const __m128i m2 = _mm_set_epi16(0, 0, 0, 0, 0, 0, -1, -1);
        const __m128i m4 = _mm_set_epi16(0, 0, 0, 0, -1, -1, -1, -1);
        std::int32_t max_score = std::numeric_limits<std::int32_t>::min();
        std::uint32_t max_i = 0, max_j = 0;
        for (std::uint32_t r = 0; r < nodes.size(); ++r) {
            const Node& node = nodes[rank_to_node_id[r]];
            const std::int16_t* prof = &profile[static_cast<std::uint64_t>(node.code) * P];
            std::int16_t* Hr = H + static_cast<std::uint64_t>(r + 1) * P;
            std::int16_t first;
            if (node.in_edges.empty()) {
                first = g_;
            } else {
                std::int16_t best = std::numeric_limits<std::int16_t>::min();
                for (std::uint32_t e : node.in_edges)
                    best = std::max(best, H[static_cast<std::uint64_t>(node_id_to_rank_[edges[e].begin_node_id] + 1) * P]);
                first = static_cast<std::int16_t>(best + g_);
            }
            const std::uint32_t npred = node.in_edges.empty() ? 1 : static_cast<std::uint32_t>(node.in_edges.size());
            __m128i carry = vneg;  // H[r][j-1] of the previous vector's last lane, broadcast
            for (std::uint32_t j = 0; j < P - 8; j += 8) {
                __m128i t = vneg;
                const __m128i pr = _mm_loadu_si128(reinterpret_cast<const __m128i*>(prof + j));
                for (std::uint32_t p = 0; p < npred; ++p) {
                    std::uint32_t pred_i = node.in_edges.empty() ? 0 : node_id_to_rank_[edges[node.in_edges[p]].begin_node_id] + 1;
                    const std::int16_t* Hp = H + static_cast<std::uint64_t>(pred_i) * P;
                    __m128i d = _mm_adds_epi16(_mm_loadu_si128(reinterpret_cast<const __m128i*>(Hp + j) - 0) , _mm_setzero_si128());
                    (void)d;
                    __m128i diag = _mm_adds_epi16(_mm_loadu_si128(reinterpret_cast<const __m128i*>(Hp + j - 1)), pr);
                    __m128i vert = _mm_adds_epi16(_mm_loadu_si128(reinterpret_cast<const __m128i*>(Hp + j)), vg);
                    t = _mm_max_epi16(t, _mm_max_epi16(diag, vert));
                }
                if (j == 0) t = _mm_insert_epi16(t, first, 0);  // column 0 has no diagonal/sequence base
                // in-vector prefix: x[k] = max_{q<=k} t[q] + (k-q)*g
                __m128i x = t;
                __m128i s = _mm_or_si128(_mm_slli_si128(x, 2), _mm_and_si128(m1, vneg));
                x = _mm_max_epi16(x, _mm_adds_epi16(s, vg));
                s = _mm_or_si128(_mm_slli_si128(x, 4), _mm_and_si128(m2, vneg));
                x = _mm_max_epi16(x, _mm_adds_epi16(s, vg2));
                s = _mm_or_si128(_mm_slli_si128(x, 8), _mm_and_si128(m4, vneg));
                x = _mm_max_epi16(x, _mm_adds_epi16(s, vg4));
                x = _mm_max_epi16(x, _mm_adds_epi16(carry, vramp));
                _mm_storeu_si128(reinterpret_cast<__m128i*>(Hr + j), x);
                carry = _mm_set1_epi16(static_cast<std::int16_t>(_mm_extract_epi16(x, 7)));
            }
            Hr[-1] = kNegInf;
            if (node.out_edges.empty() && max_score < Hr[W - 1]) { max_score = Hr[W - 1]; max_i = r + 1; max_j = W - 1; }
        }
        H[-1] = kNegInf;
        // compact view for the shared backtrack: it indexes H[i*W + j], so pass pitch P instead of W
        return backtrack<std::int16_t>(H, P, max_i, max_j, seq_codes, nodes, edges, rank_to_node_id);
    }
#endif

    AlignmentType type_;
    std::int8_t m_, n_, g_;
    bool use_simd_;
    bool last_int16_ = false;
    std::uint64_t cells_;
    std::vector<std::uint32_t> node_id_to_rank_;
    std::vector<std::int32_t> H32_;
    std::vector<std::int16_t> H16_;
};

inline std::unique_ptr<AlignmentEngine> createAlignmentEngine(AlignmentType type, std::int8_t m, std::int8_t n, std::int8_t g) {
    return std::unique_ptr<AlignmentEngine>(new AlignmentEngine(type, m, n, g));
}
inline std::unique_ptr<Graph> createGraph() { return std::unique_ptr<Graph>(new Graph()); }

}  // namespace spoa
#endif  // SPOA_RESTATED_HPP
