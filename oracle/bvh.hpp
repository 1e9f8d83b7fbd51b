// ORACLE — TEST INFRASTRUCTURE ONLY (see math.hpp header).
// Restatement of parry3d's `Bvh` (src/partitioning/bvh/*): node layout, binned / PLOC builds,
// DFS-relayout refit, ordered best-first traversal, leaf iteration, self BVTT, two-tree leaf pairs.
#pragma once
#include "math.hpp"
#include <vector>
#include <functional>
#include <cstring>
#include <cassert>
#include <numeric>

namespace pb2o {

// Rust `f32 as usize`: saturating, NaN -> 0.
static inline size_t f2usize(Real f) {
    if (!(f == f)) return 0;
    if (f <= 0.0f) return 0;
    if (f >= 18446744073709551615.0f) return SIZE_MAX;
    return (size_t)f;
}
static inline uint32_t d2u32(double f) {  // Rust `f64 as u32`
    if (!(f == f)) return 0;
    if (f <= 0.0) return 0;
    if (f >= 4294967295.0) return 0xffffffffu;
    return (uint32_t)f;
}

// bvh_tree.rs:156-205
static const uint32_t CHANGED = 0b01, CHANGE_PENDING = 0b11;
struct BvhNodeData {
    uint32_t v;
    uint32_t leaf_count() const { return v & 0x3fffffffu; }
    bool is_changed() const { return (v >> 30) == CHANGED; }
    bool is_change_pending() const { return (v >> 30) == CHANGE_PENDING; }
    void set_change_pending() { v |= CHANGE_PENDING << 30; }
    void resolve_pending_change() {
        if (is_change_pending()) v = (v & 0x3fffffffu) | (CHANGED << 30);
        else v = v & 0x3fffffffu;
    }
    BvhNodeData merged(BvhNodeData o) const {
        uint32_t lc = leaf_count() + o.leaf_count();
        uint32_t ch = (v >> 30) | (o.v >> 30);
        return BvhNodeData{lc | (ch << 30)};
    }
};

// bvh_tree.rs:456-467 — exactly 32 bytes: mins, children, maxs, data.
struct BvhNode {
    Vec3 mins;
    uint32_t children;
    Vec3 maxs;
    BvhNodeData data;
    static BvhNode zeros() { BvhNode n; n.mins = Vec3(); n.children = 0; n.maxs = Vec3(); n.data.v = 0; return n; }
    static BvhNode leaf(const Aabb& a, uint32_t id) {  // :520-527
        BvhNode n; n.mins = a.mins; n.maxs = a.maxs; n.children = id; n.data.v = 1u | (CHANGE_PENDING << 30); return n;
    }
    bool is_leaf() const { return data.leaf_count() == 1; }
    uint32_t leaf_count() const { return data.leaf_count(); }
    bool is_changed() const { return data.is_changed(); }
    Aabb aabb() const { return Aabb(mins, maxs); }
    Vec3 center() const { return pb2o::center(mins, maxs); }
    BvhNode merged(const BvhNode& o, uint32_t ch) const {  // :610-618
        BvhNode n; n.mins = vinf(mins, o.mins); n.children = ch; n.maxs = vsup(maxs, o.maxs); n.data = data.merged(o.data); return n;
    }
    bool intersects(const BvhNode& o) const { return aabb().intersects(o.aabb()); }  // :955-957
    bool contains_aabb(const Aabb& o) const { return aabb().contains(o); }
};
static_assert(sizeof(BvhNode) == 32, "BvhNode must be 32 bytes");

struct alignas(64) BvhNodeWide {
    BvhNode left, right;
    static BvhNodeWide zeros() { BvhNodeWide w; w.left = BvhNode::zeros(); w.right = BvhNode::zeros(); return w; }
    BvhNode merged(uint32_t my_id) const { return left.merged(right, my_id); }
    uint32_t leaf_count() const { return left.leaf_count() + right.leaf_count(); }
};
static_assert(sizeof(BvhNodeWide) == 64, "BvhNodeWide must be 64 bytes");

// bvh_tree.rs:1284-1420: (node id << 1) | is_right
struct BvhNodeIndex {
    size_t v = 0;
    static BvhNodeIndex left(uint32_t id) { return BvhNodeIndex{(size_t)id << 1}; }
    static BvhNodeIndex right(uint32_t id) { return BvhNodeIndex{((size_t)id << 1) | 1}; }
};

enum BuildStrategy { BINNED = 0, PLOC = 1 };

// utils/morton.rs:12-40
static inline uint64_t split_by_3_u64(uint32_t a) {
    uint64_t x = (uint64_t)a & 0x1fffff;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}
static inline uint64_t morton_encode_u64_unorm(double px, double py, double pz) {
    double s = (double)(1 << 21);
    return split_by_3_u64(d2u32(px * s)) | split_by_3_u64(d2u32(py * s)) << 1 | split_by_3_u64(d2u32(pz * s)) << 2;
}

struct Bvh {
    std::vector<BvhNodeWide> nodes;
    std::vector<BvhNodeIndex> parents;
    std::vector<BvhNodeIndex> leaf_node_indices;  // VecMap keyed by leaf id (dense here)

    BvhNode& at(BvhNodeIndex i) { return (i.v & 1) ? nodes[i.v >> 1].right : nodes[i.v >> 1].left; }

    // bvh_tree.rs:1835 / 1891-1955
    static Bvh from_leaves(BuildStrategy strategy, const Aabb* aabbs, size_t n) {
        Bvh r;
        std::vector<BvhNode> leaves;
        leaves.reserve(n);
        r.leaf_node_indices.resize(n);
        for (size_t i = 0; i < n; ++i) leaves.push_back(BvhNode::leaf(aabbs[i], (uint32_t)i));
        if (n == 0) {
        } else if (n == 1) {
            BvhNodeWide w; w.left = leaves[0]; w.right = BvhNode::zeros();
            r.nodes.push_back(w); r.parents.push_back(BvhNodeIndex());
            r.leaf_node_indices[0] = BvhNodeIndex::left(0);
        } else if (n == 2) {
            BvhNodeWide w; w.left = leaves[0]; w.right = leaves[1];
            r.nodes.push_back(w); r.parents.push_back(BvhNodeIndex());
            r.leaf_node_indices[0] = BvhNodeIndex::left(0);
            r.leaf_node_indices[1] = BvhNodeIndex::right(0);
        } else {
            r.nodes.reserve(n); r.parents.reserve(n);
            r.nodes.push_back(BvhNodeWide::zeros());
            r.parents.push_back(BvhNodeIndex());
            if (strategy == PLOC) r.rebuild_range_ploc(0, leaves);
            else r.rebuild_range_binned(0, leaves.data(), leaves.size());
            r.refit();
        }
        return r;
    }

    // bvh_binned_build.rs:11-36: same leaves (copied verbatim, change flags included), new internal nodes; nothing is refitted
    // or resolved here.
    void rebuild(BuildStrategy strategy) {
        if (nodes.size() < 2) return;
        std::vector<BvhNode> leaves;
        for (const BvhNodeWide& n : nodes) {
            if (n.left.is_leaf()) leaves.push_back(n.left);
            if (n.right.is_leaf()) leaves.push_back(n.right);
        }
        nodes.clear(); parents.clear();
        nodes.push_back(BvhNodeWide::zeros());
        parents.push_back(BvhNodeIndex());
        if (strategy == PLOC) rebuild_range_ploc(0, leaves);
        else rebuild_range_binned(0, leaves.data(), leaves.size());
    }

    // bvh_binned_build.rs:39-176
    void rebuild_range_binned(uint32_t target, BvhNode* leaves, size_t len) {
        const size_t NUM_BINS = 8;
        const Real BIN_EPSILON = 1.0e-5f;
        struct Bin { Aabb aabb; uint32_t leaf_count; };
        Bin bins[NUM_BINS];
        for (auto& b : bins) { b.aabb = Aabb::new_invalid(); b.leaf_count = 0; }
        assert(len > 1);

        // Aabb::from_points(centers) — local_point_cloud_aabb (aabb_utils.rs:97-115)
        Aabb caabb; caabb.mins = caabb.maxs = leaves[0].center();
        for (size_t i = 1; i < len; ++i) { Vec3 c = leaves[i].center(); caabb.mins = vinf(caabb.mins, c); caabb.maxs = vsup(caabb.maxs, c); }
        int axis = imax(caabb.extents());
        Real r0 = caabb.mins[axis], r1 = caabb.maxs[axis];
        Real k1 = (Real)NUM_BINS * (1.0f - BIN_EPSILON) / (r1 - r0);
        Real k0 = r0;
        for (size_t i = 0; i < len; ++i) {
            size_t b = f2usize(k1 * (leaves[i].center()[axis] - k0));
            bins[b].aabb.merge(leaves[i].aabb());
            bins[b].leaf_count += 1;
        }
        Bin right_merges[NUM_BINS];
        for (size_t i = 0; i < NUM_BINS; ++i) right_merges[i] = bins[i];
        Bin right_acc = bins[NUM_BINS - 1];
        for (size_t i = 1; i < NUM_BINS - 1; ++i) {
            right_acc.aabb.merge(right_merges[NUM_BINS - 1 - i].aabb);
            right_acc.leaf_count += right_merges[NUM_BINS - 1 - i].leaf_count;
            right_merges[NUM_BINS - 1 - i] = right_acc;
        }
        Real best_cost = REAL_MAX;
        size_t best_plane = 0;
        Bin left_merge = bins[0];
        uint32_t best_leaf_count = bins[0].leaf_count;
        for (size_t i = 0; i < NUM_BINS - 1; ++i) {
            const Bin& right = right_merges[i + 1];
            Real cost = left_merge.aabb.volume() * (Real)left_merge.leaf_count + right.aabb.volume() * (Real)right.leaf_count;
            if (cost < best_cost) { best_cost = cost; best_plane = i; best_leaf_count = left_merge.leaf_count; }
            left_merge.aabb.merge(bins[i + 1].aabb);
            left_merge.leaf_count += bins[i + 1].leaf_count;
        }
        size_t mid = best_leaf_count;
        if (mid == 0 || mid == len) {
            mid = len / 2;
        } else {
            auto bin = [&](size_t id) { return f2usize(k1 * (leaves[id].center()[axis] - k0)); };
            size_t left_id = 0, right_id = mid;
            while (left_id != mid && right_id != len) {
                bool brk = false;
                while (bin(left_id) <= best_plane) { left_id++; if (left_id == mid) { brk = true; break; } }
                if (brk) break;
                while (bin(right_id) > best_plane) { right_id++; if (right_id == len) { brk = true; break; } }
                if (brk) break;
                std::swap(leaves[left_id], leaves[right_id]);
                left_id++; right_id++;
            }
        }
        BvhNode* left_leaves = leaves; size_t nl = mid;
        BvhNode* right_leaves = leaves + mid; size_t nr = len - mid;
        assert(nl > 0 && nr > 0);
        if (nl == 1) {
            nodes[target].left = left_leaves[0];
            if (nodes[target].left.is_leaf()) leaf_node_indices[nodes[target].left.children] = BvhNodeIndex::left(target);
            else parents[nodes[target].left.children] = BvhNodeIndex::left(target);
        } else {
            uint32_t lid = (uint32_t)nodes.size();
            nodes.push_back(BvhNodeWide::zeros());
            parents.push_back(BvhNodeIndex::left(target));
            rebuild_range_binned(lid, left_leaves, nl);
            nodes[target].left = nodes[lid].merged(lid);
        }
        if (nr == 1) {
            nodes[target].right = right_leaves[0];
            if (nodes[target].right.is_leaf()) leaf_node_indices[nodes[target].right.children] = BvhNodeIndex::right(target);
            else parents[nodes[target].right.children] = BvhNodeIndex::right(target);
        } else {
            uint32_t rid = (uint32_t)nodes.size();
            nodes.push_back(BvhNodeWide::zeros());
            parents.push_back(BvhNodeIndex::right(target));
            rebuild_range_binned(rid, right_leaves, nr);
            nodes[target].right = nodes[rid].merged(rid);
        }
    }

    // bvh_ploc_build.rs:10-94
    void rebuild_range_ploc(uint32_t target, std::vector<BvhNode>& leaves) {
        Aabb aabb; aabb.mins = aabb.maxs = leaves[0].center();
        for (size_t i = 1; i < leaves.size(); ++i) { Vec3 c = leaves[i].center(); aabb.mins = vinf(aabb.mins, c); aabb.maxs = vsup(aabb.maxs, c); }
        Vec3 e = aabb.extents();
        Vec3 inv(1.0f / e.x, 1.0f / e.y, 1.0f / e.z);
        // sort_by_cached_key: stable sort by key
        std::vector<uint64_t> keys(leaves.size());
        for (size_t i = 0; i < leaves.size(); ++i) {
            Vec3 d = leaves[i].center() - aabb.mins;
            Vec3 c(d.x * inv.x, d.y * inv.y, d.z * inv.z);
            keys[i] = morton_encode_u64_unorm((double)c.x, (double)c.y, (double)c.z);
        }
        std::vector<size_t> order(leaves.size());
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return keys[a] < keys[b]; });
        {
            std::vector<BvhNode> sorted(leaves.size());
            for (size_t i = 0; i < order.size(); ++i) sorted[i] = leaves[order[i]];
            leaves.swap(sorted);
        }
        const size_t SEARCH_RADIUS = 16;
        std::vector<size_t> cand(leaves.size(), SIZE_MAX);
        std::vector<BvhNode> next;
        next.reserve(leaves.size());
        while (leaves.size() > 1) {
            size_t n = leaves.size();
            for (size_t i = 0; i < n; ++i) {
                Real best_sah = REAL_MAX; size_t best = SIZE_MAX;
                size_t lo = i >= SEARCH_RADIUS ? i - SEARCH_RADIUS : 0;
                size_t hi = std::min(i + SEARCH_RADIUS, n - 1);
                for (size_t k = lo; k <= hi; ++k) {
                    if (k != i) {
                        Real sah = leaves[i].aabb().merged(leaves[k].aabb()).half_area();
                        if (sah < best_sah) { best_sah = sah; best = k; }
                    }
                }
                cand[i] = best;
            }
            for (size_t i = 0; i < n; ++i) {
                size_t k = cand[i];
                if (cand[k] == i) {
                    if (i > k) continue;
                    BvhNode left = leaves[i], right = leaves[k];
                    BvhNodeWide wide; wide.left = left; wide.right = right;
                    uint32_t id;
                    if (n == 2) { nodes[target] = wide; id = target; }
                    else {
                        id = (uint32_t)nodes.size();
                        BvhNode parent = wide.merged(id);
                        nodes.push_back(wide);
                        parents.push_back(BvhNodeIndex());
                        next.push_back(parent);
                    }
                    if (left.is_leaf()) leaf_node_indices[left.children] = BvhNodeIndex::left(id);
                    else parents[left.children] = BvhNodeIndex::left(id);
                    if (right.is_leaf()) leaf_node_indices[right.children] = BvhNodeIndex::right(id);
                    else parents[right.children] = BvhNodeIndex::right(id);
                } else {
                    next.push_back(leaves[i]);
                }
            }
            leaves.swap(next);
            next.clear();
        }
    }

    // bvh_refit.rs:170-257
    void refit() {
        std::vector<BvhNodeWide> target;
        if (nodes.empty()) { parents.clear(); }
        else if (nodes[0].leaf_count() <= 2) {
            target.push_back(nodes[0]);
            target[0].left.data.resolve_pending_change();
            if (target[0].right.leaf_count() > 0) target[0].right.data.resolve_pending_change();
            parents.clear(); parents.push_back(BvhNodeIndex());
        } else {
            target.assign(nodes.size(), BvhNodeWide::zeros());
            uint32_t len = 1;
            uint32_t lc = nodes[0].left.children, rc = nodes[0].right.children;
            if (!nodes[0].left.is_leaf()) refit_recurse(target, lc, len, BvhNodeIndex::left(0));
            else { target[0].left = nodes[0].left; target[0].left.data.resolve_pending_change(); }
            if (!nodes[0].right.is_leaf()) refit_recurse(target, rc, len, BvhNodeIndex::right(0));
            else { target[0].right = nodes[0].right; target[0].right.data.resolve_pending_change(); }
            target.resize(len); parents.resize(len);
        }
        nodes.swap(target);
    }
    // bvh_refit.rs:259-320
    void refit_recurse(std::vector<BvhNodeWide>& target, uint32_t source_id, uint32_t& next_id, BvhNodeIndex parent) {
        uint32_t target_id = next_id++;
        const BvhNodeWide& node = nodes[source_id];
        if (!node.left.is_leaf()) refit_recurse(target, node.left.children, next_id, BvhNodeIndex::left(target_id));
        else {
            target[target_id].left = node.left;
            target[target_id].left.data.resolve_pending_change();
            leaf_node_indices[node.left.children] = BvhNodeIndex::left(target_id);
        }
        if (!node.right.is_leaf()) refit_recurse(target, node.right.children, next_id, BvhNodeIndex::right(target_id));
        else {
            target[target_id].right = node.right;
            target[target_id].right.data.resolve_pending_change();
            leaf_node_indices[node.right.children] = BvhNodeIndex::right(target_id);
        }
        BvhNode merged = target[target_id].left.merged(target[target_id].right, target_id);
        ((parent.v & 1) ? target[parent.v >> 1].right : target[parent.v >> 1].left) = merged;
        parents[target_id] = parent;
    }

    // bvh_refit.rs:326-375
    void refit_without_opt() {
        if (leaf_count() > 2) {
            BvhNodeWide root = nodes[0];
            if (!root.left.is_leaf()) recurse_refit_without_opt(root.left.children, BvhNodeIndex::left(0));
            if (!root.right.is_leaf()) recurse_refit_without_opt(root.right.children, BvhNodeIndex::right(0));
        }
    }
    void recurse_refit_without_opt(uint32_t id, BvhNodeIndex parent) {
        bool ll = nodes[id].left.is_leaf(), rl = nodes[id].right.is_leaf();
        uint32_t lc = nodes[id].left.children, rc = nodes[id].right.children;
        if (!ll) recurse_refit_without_opt(lc, BvhNodeIndex::left(id)); else nodes[id].left.data.resolve_pending_change();
        if (!rl) recurse_refit_without_opt(rc, BvhNodeIndex::right(id)); else nodes[id].right.data.resolve_pending_change();
        at(parent) = nodes[id].left.merged(nodes[id].right, id);
    }

    uint32_t leaf_count() const { return nodes.empty() ? 0 : nodes[0].leaf_count(); }  // bvh_tree.rs:2295

    // bvh_insert.rs:209-231 (existing leaves only; inserting unknown ids is out of scope)
    void insert_or_update_partially(const Aabb& aabb, uint32_t leaf, Real margin) {
        BvhNode& node = at(leaf_node_indices[leaf]);
        if (margin > 0.0f) {
            if (!node.contains_aabb(aabb)) {
                node.mins = aabb.mins - Vec3(margin, margin, margin);
                node.maxs = aabb.maxs + Vec3(margin, margin, margin);
                node.data.set_change_pending();
            }
        } else { node.mins = aabb.mins; node.maxs = aabb.maxs; }
    }

    // bvh_traverse.rs:335-417. L must expose .cost(). Returns false for None.
    template <class L, class AabbCost, class LeafCost>
    bool find_best(Real max_cost, AabbCost aabb_cost, LeafCost leaf_cost, uint32_t& out_id, L& out_val) const {
        std::vector<uint32_t> stack;
        stack.reserve(32);
        bool have_best = false;
        L best_val{};
        Real best_cost = max_cost;
        uint32_t best_id = UINT32_MAX;
        uint32_t curr_id = 0;
        if (nodes.empty()) return false;
        if (nodes[0].right.leaf_count() == 0) {
            const BvhNode& leaf = nodes[0].left;
            if (aabb_cost(leaf, max_cost) < max_cost) {
                L cost;
                if (!leaf_cost(leaf.children, best_cost, cost)) return false;
                if (cost.cost() < max_cost) { out_id = leaf.children; out_val = cost; return true; }
                return false;
            }
            return false;
        }
        for (;;) {
            const BvhNodeWide& node = nodes[curr_id];
            const BvhNode* left = &node.left;
            const BvhNode* right = &node.right;
            Real left_score = aabb_cost(*left, best_cost);
            Real right_score = aabb_cost(*right, best_cost);
            if (left_score > right_score) { std::swap(left_score, right_score); std::swap(left, right); }
            bool found_next = false;
            if (left_score < best_cost && left_score != REAL_MAX) {
                if (left->is_leaf()) {
                    L v;
                    if (leaf_cost(left->children, best_cost, v)) {
                        Real s = v.cost();
                        if (s < best_cost) { best_val = v; have_best = true; best_cost = s; best_id = left->children; }
                    }
                } else { curr_id = left->children; found_next = true; }
            }
            if (right_score < best_cost && right_score != REAL_MAX) {
                if (right->is_leaf()) {
                    L v;
                    if (leaf_cost(right->children, best_cost, v)) {
                        Real s = v.cost();
                        if (s < best_cost) { best_val = v; have_best = true; best_cost = s; best_id = right->children; }
                    }
                } else if (found_next) stack.push_back(right->children);
                else { curr_id = right->children; found_next = true; }
            }
            if (!found_next) {
                if (!stack.empty()) { curr_id = stack.back(); stack.pop_back(); }
                else {
                    if (have_best) { out_id = best_id; out_val = best_val; }
                    return have_best;
                }
            }
        }
    }

    // bvh_traverse.rs:8-70 `Leaves` iterator, drained into `out` in iteration order.
    template <class Check>
    void leaves(Check check, std::vector<uint32_t>& out) const {
        std::vector<const BvhNode*> stack;
        const BvhNode* next = nullptr;
        if (!nodes.empty()) {
            const BvhNodeWide& root = nodes[0];
            if (check(root.left)) next = &root.left;
            if (root.right.leaf_count() > 0 && check(root.right)) stack.push_back(&root.right);
        }
        for (;;) {
            if (!next) { if (stack.empty()) return; next = stack.back(); stack.pop_back(); }
            const BvhNode* node = next; next = nullptr;
            if (node->is_leaf()) { out.push_back(node->children); continue; }
            const BvhNodeWide& ch = nodes[node->children];
            if (check(ch.left)) next = &ch.left;
            if (check(ch.right)) { if (!next) next = &ch.right; else stack.push_back(&ch.right); }
        }
    }
    // bvh_queries.rs:203-205
    void intersect_aabb(const Aabb& q, std::vector<uint32_t>& out) const {
        leaves([&](const BvhNode& n) { return n.aabb().intersects(q); }, out);
    }

    // bvh_traverse_bvtt.rs:19-204
    template <bool CD, class F>
    void traverse_bvtt_single_tree(F& f) const {
        if (nodes.empty() || nodes[0].right.leaf_count() == 0) return;
        std::vector<uint32_t> stack;
        self_intersect_node<CD>(stack, 0, f);
    }
    template <bool CD, class F>
    void self_intersect_node(std::vector<uint32_t>& stack, uint32_t id, F& f) const {
        const BvhNodeWide& node = nodes[id];
        if (CD && !node.right.is_changed() && !node.left.is_changed()) return;
        bool lr = node.left.intersects(node.right);
        uint32_t lc = node.left.children, rc = node.right.children;
        bool ll = node.left.is_leaf(), rl = node.right.is_leaf();
        if ((!CD || node.left.is_changed()) && !ll) self_intersect_node<CD>(stack, lc, f);
        if ((!CD || node.right.is_changed()) && !rl) self_intersect_node<CD>(stack, rc, f);
        if (lr) {
            if (ll && rl) f(lc, rc);
            else if (ll && !rl) traverse_single_subtree<CD>(stack, node.left, rc, f);
            else if (!ll && rl) traverse_single_subtree<CD>(stack, node.right, lc, f);
            else traverse_two_branches<CD>(stack, lc, rc, f);
        }
    }
    template <bool CD, class F>
    void dispatch_pair(std::vector<uint32_t>& stack, bool check, const BvhNode& a, const BvhNode& b, F& f) const {
        if (!check) return;
        bool al = a.is_leaf(), bl = b.is_leaf();
        if (al && bl) f(a.children, b.children);
        else if (al && !bl) traverse_single_subtree<CD>(stack, a, b.children, f);
        else if (!al && bl) traverse_single_subtree<CD>(stack, b, a.children, f);
        else traverse_two_branches<CD>(stack, a.children, b.children, f);
    }
    template <bool CD, class F>
    void traverse_two_branches(std::vector<uint32_t>& stack, uint32_t a, uint32_t b, F& f) const {
        const BvhNode& l1 = nodes[a].left; const BvhNode& r1 = nodes[a].right;
        const BvhNode& l2 = nodes[b].left; const BvhNode& r2 = nodes[b].right;
        bool ll = (!CD || l1.is_changed() || l2.is_changed()) && l1.intersects(l2);
        bool lr = (!CD || l1.is_changed() || r2.is_changed()) && l1.intersects(r2);
        bool rl = (!CD || r1.is_changed() || l2.is_changed()) && r1.intersects(l2);
        bool rr = (!CD || r1.is_changed() || r2.is_changed()) && r1.intersects(r2);
        dispatch_pair<CD>(stack, ll, l1, l2, f);
        dispatch_pair<CD>(stack, lr, l1, r2, f);
        dispatch_pair<CD>(stack, rl, r1, l2, f);
        dispatch_pair<CD>(stack, rr, r1, r2, f);
    }
    template <bool CD, class F>
    void traverse_single_subtree(std::vector<uint32_t>& stack, const BvhNode& node, uint32_t subtree, F& f) const {
        uint32_t curr = subtree;
        bool node_changed = node.is_changed();
        for (;;) {
            const BvhNode& left = nodes[curr].left; const BvhNode& right = nodes[curr].right;
            bool lchk = (!CD || node_changed || left.is_changed()) && node.intersects(left);
            bool rchk = (!CD || node_changed || right.is_changed()) && node.intersects(right);
            bool found_next = false;
            if (lchk) { if (left.is_leaf()) f(node.children, left.children); else { curr = left.children; found_next = true; } }
            if (rchk) {
                if (right.is_leaf()) f(node.children, right.children);
                else if (!found_next) { curr = right.children; found_next = true; }
                else stack.push_back(right.children);
            }
            if (!found_next) { if (stack.empty()) return; curr = stack.back(); stack.pop_back(); }
        }
    }

    // bvh_traverse_bvtt.rs:210-316, drained. NOTE: root-level pairs are pushed unchecked.
    template <class Check, class F>
    void leaf_pairs(const Bvh& other, Check check, F& f) const {
        typedef std::pair<const BvhNode*, const BvhNode*> P;
        std::vector<P> stack;
        bool has_next = false; P next;
        if (!nodes.empty() && !other.nodes.empty()) {
            const BvhNodeWide& r1 = nodes[0]; const BvhNodeWide& r2 = other.nodes[0];
            if (r1.left.leaf_count() > 0 && r2.right.leaf_count() > 0) stack.push_back(P(&r1.left, &r2.right));
            if (r1.right.leaf_count() > 0) {
                if (r2.right.leaf_count() > 0) stack.push_back(P(&r1.right, &r2.right));
                stack.push_back(P(&r1.right, &r2.left));
            }
            next = P(&r1.left, &r2.left); has_next = true;
        }
        for (;;) {
            if (!has_next) { if (stack.empty()) return; next = stack.back(); stack.pop_back(); has_next = true; }
            const BvhNode* n1 = next.first; const BvhNode* n2 = next.second; has_next = false;
            bool l1 = n1->is_leaf(), l2 = n2->is_leaf();
            if (l1 && l2) { f(n1->children, n2->children); }
            else if (l1 && !l2) {
                const BvhNodeWide& c2 = other.nodes[n2->children];
                if (check(*n1, c2.left)) { next = P(n1, &c2.left); has_next = true; }
                if (check(*n1, c2.right)) { if (!has_next) { next = P(n1, &c2.right); has_next = true; } else stack.push_back(P(n1, &c2.right)); }
            } else if (!l1 && l2) {
                const BvhNodeWide& c1 = nodes[n1->children];
                if (check(c1.left, *n2)) { next = P(&c1.left, n2); has_next = true; }
                if (check(c1.right, *n2)) { if (!has_next) { next = P(&c1.right, n2); has_next = true; } else stack.push_back(P(&c1.right, n2)); }
            } else {
                const BvhNodeWide& c1 = nodes[n1->children]; const BvhNodeWide& c2 = other.nodes[n2->children];
                if (check(c1.left, c2.left)) stack.push_back(P(&c1.left, &c2.left));
                if (check(c1.right, c2.left)) stack.push_back(P(&c1.right, &c2.left));
                if (check(c1.left, c2.right)) stack.push_back(P(&c1.left, &c2.right));
                if (check(c1.right, c2.right)) stack.push_back(P(&c1.right, &c2.right));
            }
        }
    }

};

}  // namespace pb2o
