// Host-side forest flattener (no CUDA calls): one fitted scikit-learn tree ->
// 8-byte packed nodes in level order (sibling pairs) + the leaf -> in-bag member lists that
// replace the dense (T,N) y_train_leaves_ / y_weights_ arrays of the reference
// (skgarden/quantile/ensemble.py:83-101).  Layout documented in include/qforest_b200.h.
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <vector>

#include "qf_internal.h"

namespace {

// Largest float32 <= t.  sklearn compares the float32 feature with the float64
// threshold in double (sklearn:tree/_tree.pyx:989); for float32 x,
//   (double)x <= t   <=>   x <= floor_f32(t)
// so the device can compare in float32 and still decide exactly like sklearn.
inline float floor_to_f32(double t) {
    float f = static_cast<float>(t);            // round to nearest
    if (static_cast<double>(f) > t) f = std::nextafterf(f, -INFINITY);
    return f;
}

inline uint32_t f32_bits(float f) {
    uint32_t u;
    std::memcpy(&u, &f, sizeof(u));
    return u;
}

}  // namespace

extern "C" int64_t qf_count_inbag(int64_t n_train, const int64_t* counts) {
    if (counts == nullptr) return n_train;
    int64_t m = 0;
    for (int64_t j = 0; j < n_train; ++j) m += counts[j] > 0;
    return m;
}

extern "C" int qf_pack_tree(int64_t n_nodes, const int64_t* left, const int64_t* right,
                            const int64_t* feature, const double* threshold,
                            int32_t feature_bits, int64_t n_train, const int64_t* train_leaves,
                            const int64_t* counts, const int64_t* sorter, uint64_t member_base,
                            uint64_t* nodes_out, int32_t* node_id_out, uint64_t* members_out,
                            int64_t* n_inbag_out, int32_t* max_leaf_members_out,
                            int32_t* max_depth_out, double* sum_sq_leaf_members_out) {
    // train_leaves == NULL: structure only (nodes, node ids, depth); the leaves stay empty
    // and the member lists are built on the device (qf_build_csr)
    const bool structure_only = train_leaves == nullptr;
    if (n_nodes <= 0 || !left || !right || !feature || !threshold || !nodes_out || !node_id_out ||
        (!structure_only && (!sorter || !members_out || n_train <= 0)))
        return qf::fail(QF_ERR_INVALID, "qf_pack_tree: null or empty argument");
    if (feature_bits < 1 || feature_bits > 24)
        return qf::fail(QF_ERR_INVALID, "qf_pack_tree: feature_bits must be in [1,24]");
    if (n_nodes >= (int64_t(1) << 31) || (!structure_only && n_train >= (int64_t(1) << 31)))
        return qf::fail(QF_ERR_LIMIT, "qf_pack_tree: more than 2^31 nodes or training rows");
    const uint32_t max_off = (1u << (31 - feature_bits)) - 1u;

    // pass 1: level-order numbering with sibling pairs.  Slot 0 = root, slot 1 = padding (so
    // that every pair starts on a 16-byte boundary), then level after level, left to right, the
    // two children of every internal node side by side.  Rows that are neighbours in feature
    // space sit on neighbouring nodes of a level, so the 32 lanes of a leaf-lookup warp touch
    // few 32-byte sectors per visit.
    const int64_t n_slots = n_nodes + 1;
    std::vector<int32_t> pos_of(n_nodes, -1);
    std::vector<int32_t> depth_of(n_nodes, 0);
    pos_of[0] = 0;
    node_id_out[0] = 0;
    node_id_out[1] = -1;
    nodes_out[1] = 0;                                      // never reached by a walk, not a leaf
    int32_t next = 2, max_depth = 0;
    // Generalisation used for experiments: QF_NODE_BLOCK_LEVELS=h groups h levels of descendants of
    // a node into one contiguous block (blocks in level order); unset / 0 = one block = plain level order.
    int block_levels = INT32_MAX;
    if (const char* env = std::getenv("QF_NODE_BLOCK_LEVELS")) {
        const int h = std::atoi(env);
        if (h > 0) block_levels = h;
    }
    std::vector<int64_t> roots(1, 0), next_roots, level, next_level;
    while (!roots.empty()) {
        next_roots.clear();
        for (const int64_t r : roots) {
            level.assign(1, r);
            for (int k = 0; k < block_levels && !level.empty(); ++k) {
                next_level.clear();
                for (const int64_t nd : level) {
                    if (left[nd] == -1) continue;          // sklearn: leaf <=> children_left == -1 (_tree.pyx:978)
                    const int64_t kids[2] = {left[nd], right[nd]};
                    for (const int64_t c : kids) {
                        if (c <= 0 || c >= n_nodes || pos_of[c] != -1 || next >= n_slots)
                            return qf::fail(QF_ERR_INVALID, "qf_pack_tree: children arrays do not form a tree");
                        pos_of[c] = next;
                        node_id_out[next] = static_cast<int32_t>(c);
                        ++next;
                        depth_of[c] = depth_of[nd] + 1;
                        if (depth_of[c] > max_depth) max_depth = depth_of[c];
                        next_level.push_back(c);
                    }
                }
                level.swap(next_level);
            }
            for (const int64_t nd : level)
                if (left[nd] != -1) next_roots.push_back(nd);
        }
        roots.swap(next_roots);
    }
    if (next != n_slots)
        return qf::fail(QF_ERR_INVALID, "qf_pack_tree: unreachable nodes in tree");

    // the leaves from left to right (depth-first, left subtree first): the order of the member
    // lists, and of the leaf ordinals a structure-only packing leaves in lo32
    std::vector<int64_t> leaf_seq;
    leaf_seq.reserve(n_nodes / 2 + 1);
    {
        std::vector<int64_t> stack;
        stack.reserve(128);
        stack.push_back(0);
        while (!stack.empty()) {
            const int64_t nd = stack.back();
            stack.pop_back();
            if (left[nd] == -1) {
                leaf_seq.push_back(nd);
            } else {
                stack.push_back(right[nd]);
                stack.push_back(left[nd]);                 // popped first
            }
        }
    }

    // pass 2: per-leaf in-bag statistics (ensemble.py:94-99: weights = count / sum of counts)
    std::vector<int64_t> leaf_count_sum(n_nodes, 0);
    std::vector<uint32_t> leaf_members(n_nodes, 0);
    int64_t n_inbag = 0;
    for (int64_t j = 0; !structure_only && j < n_train; ++j) {
        const int64_t c = counts ? counts[j] : 1;
        if (c <= 0) continue;                              // out-of-bag: -1 / weight 0 in the reference
        const int64_t lf = train_leaves[j];
        if (lf < 0 || lf >= n_nodes || left[lf] != -1)
            return qf::fail(QF_ERR_INVALID, "qf_pack_tree: train_leaves entry is not a leaf id");
        leaf_count_sum[lf] += c;
        leaf_members[lf] += 1;
        ++n_inbag;
    }
    if (member_base + static_cast<uint64_t>(structure_only ? n_nodes : n_inbag) > 0xFFFFFFFFull)
        return qf::fail(QF_ERR_LIMIT, "qf_pack_tree: more than 2^32-1 member entries in the forest");

    // pass 3: packed nodes.  Leaves get their member range in left-to-right leaf order (lists of
    // leaves that are close in feature space are close in memory); structure only: lo32 = member_base
    // + left-to-right ordinal of the leaf, which is what qf_build_csr keys its scan on.
    std::vector<uint32_t> cursor(n_nodes, 0);              // indexed by sklearn id
    uint64_t running = member_base;
    uint32_t max_members = 0;
    double sum_sq = 0.0;
    for (const int64_t nd : leaf_seq) {
        const uint32_t m = leaf_members[nd];
        cursor[nd] = static_cast<uint32_t>(running);
        nodes_out[pos_of[nd]] = (static_cast<uint64_t>(0x80000000u | m) << 32) | static_cast<uint32_t>(running);
        running += structure_only ? 1 : m;
        if (m > max_members) max_members = m;
        sum_sq += static_cast<double>(m) * static_cast<double>(m);
    }
    for (int64_t nd = 0; nd < n_nodes; ++nd) {
        if (left[nd] == -1) continue;
        const int64_t p = pos_of[nd];
        if (pos_of[right[nd]] != pos_of[left[nd]] + 1)
            return qf::fail(QF_ERR_INVALID, "qf_pack_tree: internal numbering error");
        const int64_t off = static_cast<int64_t>(pos_of[left[nd]]) - p;
        const int64_t f = feature[nd];
        if (off < 1 || static_cast<uint64_t>(off) > max_off)
            return qf::fail(QF_ERR_LIMIT, "qf_pack_tree: child offset does not fit 31-feature_bits bits");
        if (f < 0 || f >= (int64_t(1) << feature_bits))
            return qf::fail(QF_ERR_INVALID, "qf_pack_tree: feature index does not fit feature_bits");
        const uint32_t hi = (static_cast<uint32_t>(off) << feature_bits) | static_cast<uint32_t>(f);
        nodes_out[p] = (static_cast<uint64_t>(hi) << 32) | f32_bits(floor_to_f32(threshold[nd]));
    }

    // pass 4: members in rank order -> every leaf's list ends up sorted by rank
    for (int64_t k = 0; !structure_only && k < n_train; ++k) {
        const int64_t j = sorter[k];
        if (j < 0 || j >= n_train)
            return qf::fail(QF_ERR_INVALID, "qf_pack_tree: sorter is not a permutation");
        const int64_t c = counts ? counts[j] : 1;
        if (c <= 0) continue;
        const int64_t lf = train_leaves[j];
        // ensemble.py:98-99: int64 / int64 -> float64, stored into the float32 y_weights_
        const float w = static_cast<float>(static_cast<double>(c) / static_cast<double>(leaf_count_sum[lf]));
        const uint64_t slot = cursor[lf]++ - member_base;
        members_out[slot] = (static_cast<uint64_t>(f32_bits(w)) << 32) | static_cast<uint32_t>(k);
    }

    if (n_inbag_out) *n_inbag_out = n_inbag;
    if (max_leaf_members_out) *max_leaf_members_out = static_cast<int32_t>(max_members);
    if (max_depth_out) *max_depth_out = max_depth;
    if (sum_sq_leaf_members_out) *sum_sq_leaf_members_out = sum_sq;
    return QF_OK;
}
