// Host-side planning: the flattened species tree and the two op lists the tree-walking kernels execute.
//
//   ProgramBuilder   stack-machine program of the pruning kernel (prune.cuh): LEAVES / GEMM(+park) / ROOT
//   ScheduleBuilder  slot-machine schedule of the reconstruction kernel (pupko.cuh): SET / MUL / SPILL / FILL / ROOT
//
// Both visit the tree in the post-order the reference's recursion implies (clade::apply_reverse_level_order feeds
// compute_node_probability children before parents, src/core.cpp:133-144) and keep the reference's factor order
// wherever a re-association could change the last bit (nodes with more than two children).
#pragma once

#include <algorithm>
#include <vector>

#include "../../include/cafe_b200.h"
#include "common.cuh"

namespace cafe {

struct HostTree {
    int n_nodes = 0;
    std::vector<int> parent, child_offset, child_list, leaf_col, lambda_index;
    std::vector<double> branch;
    std::vector<long> branch_key;       // long(t * 1000)                         src/matrix_cache.h:50
    int n_internal = 0;
    int n_lambdas = 1;
    bool is_leaf(int v) const { return child_offset[v] == child_offset[v + 1]; }
};

// Copies and validates the caller's tree; returns an error text or nullptr.
inline const char* import_tree(HostTree& t, const cafe_b200_tree* tree, int n_leaves)
{
    const int nn = tree->n_nodes;
    t.n_nodes = nn;
    t.parent.assign(tree->parent, tree->parent + nn);
    t.child_offset.assign(tree->child_offset, tree->child_offset + nn + 1);
    t.child_list.assign(tree->child_list, tree->child_list + (nn - 1));
    t.leaf_col.assign(tree->leaf_col, tree->leaf_col + nn);
    t.lambda_index.assign(tree->lambda_index, tree->lambda_index + nn);
    t.branch.assign(tree->branch, tree->branch + nn);
    t.branch_key.resize(nn);
    t.n_internal = 0;
    t.n_lambdas = 1;
    if (t.child_offset[0] != 0 || t.child_offset[nn] != nn - 1) return "child_offset does not describe n_nodes-1 edges";
    int leaves = 0;
    for (int v = 0; v < nn; ++v) {
        if ((t.parent[v] < 0) != (v == nn - 1)) return "the root must be the last node and the only one without parent";
        if (v < nn - 1 && (t.parent[v] <= v || t.parent[v] >= nn)) return "children must precede their parent";
        if (t.child_offset[v + 1] < t.child_offset[v]) return "child_offset not monotone";
        for (int e = t.child_offset[v]; e < t.child_offset[v + 1]; ++e)
            if (t.child_list[e] < 0 || t.child_list[e] >= v || t.parent[t.child_list[e]] != v) return "child_list inconsistent with parent";
        if (t.is_leaf(v)) {
            ++leaves;
            if (n_leaves >= 0 && (t.leaf_col[v] < 0 || t.leaf_col[v] >= n_leaves)) return "leaf_col out of range";
        }
        else t.n_internal++;
        if (t.lambda_index[v] < 0) return "negative lambda index";
        t.n_lambdas = std::max(t.n_lambdas, t.lambda_index[v] + 1);
        t.branch_key[v] = (long)(t.branch[v] * 1000);
    }
    if (n_leaves >= 0 && leaves != n_leaves) return "n_leaves does not match the tree";
    if (t.is_leaf(nn - 1)) return "the root is a leaf";
    return nullptr;
}

// ------------------------------------------------------------------------------------------------------------------
// Pruning program.  A node's vector is the product of its children's factors in Newick order
// (src/probability.cpp:229-238: node_probs = 1; node_probs *= factor_1; *= factor_2; ...).  For nodes with at most two
// children the order is free (a * b == b * a bit for bit), so the internal child whose subtree needs the deeper stack
// goes first and leaves go last; nodes with more children keep Newick order.
//
// While a later internal child's subtree is evaluated, the product so far is PARKED on a stack; `stack` is the index
// of that entry (= the number of products parked by enclosing nodes).  depth = entries the whole tree needs
// (Sethi-Ullman: <= log2(leaves) for binary trees).
// ------------------------------------------------------------------------------------------------------------------
struct ProgOp {
    int type;        // POP_*
    int node;        // GEMM: the internal child whose edge matrix is applied; LEAVES / ROOT: the node itself
    int flags;       // PF_*
    int stack;       // stack index of the parked product this GEMM pops and/or pushes
    int leaf_begin;  // into Program::leaves
    int n_pre;
    int n_post;
};

struct Program {
    std::vector<ProgOp> ops;
    std::vector<int> leaves;    // leaf node ids, grouped per op
    int depth = 0;
    int n_gemm = 0;
};

class ProgramBuilder {
public:
    explicit ProgramBuilder(const HostTree& t) : tree(t) { compute_depth(); }

    Program build()
    {
        const int root = tree.n_nodes - 1;
        emit(root, 0);
        out.ops.push_back({POP_ROOT, root, 0, 0, (int)out.leaves.size(), 0, 0});
        return out;
    }

private:
    const HostTree& tree;
    std::vector<int> need;      // parked entries the subtree of v needs
    Program out;

    // children of v in evaluation order
    std::vector<int> order_of(int v) const
    {
        std::vector<int> order;
        for (int e = tree.child_offset[v]; e < tree.child_offset[v + 1]; ++e) order.push_back(tree.child_list[e]);
        if (order.size() <= 2)
            std::stable_sort(order.begin(), order.end(), [this](int a, int b) {
                const int na = tree.is_leaf(a) ? -1 : need[a], nb = tree.is_leaf(b) ? -1 : need[b];
                return na > nb;
            });
        return order;
    }

    void compute_depth()
    {
        need.assign(tree.n_nodes, 0);
        for (int v = 0; v < tree.n_nodes; ++v) {
            if (tree.is_leaf(v)) continue;
            int n = 0, seen = 0;
            for (int c : order_of(v)) {
                if (tree.is_leaf(c)) continue;
                n = std::max(n, need[c] + (seen > 0 ? 1 : 0));
                ++seen;
            }
            need[v] = n;
        }
    }

    void emit(int v, int level)
    {
        const std::vector<int> order = order_of(v);
        std::vector<size_t> internal;
        for (size_t i = 0; i < order.size(); ++i)
            if (!tree.is_leaf(order[i])) internal.push_back(i);
        if (internal.empty()) {
            out.ops.push_back({POP_LEAVES, v, 0, 0, (int)out.leaves.size(), (int)order.size(), 0});
            out.leaves.insert(out.leaves.end(), order.begin(), order.end());
            return;
        }
        for (size_t i = 0; i < internal.size(); ++i) {
            const int c = order[internal[i]];
            emit(c, i == 0 ? level : level + 1);
            ProgOp op = {POP_GEMM, c, 0, level, (int)out.leaves.size(), 0, 0};
            if (i > 0) op.flags |= PF_PARKED;
            if (i + 1 < internal.size()) op.flags |= PF_PARK;
            if (op.flags) out.depth = std::max(out.depth, level + 1);
            if (i == 0)
                for (size_t j = 0; j < internal[0]; ++j) { out.leaves.push_back(order[j]); op.n_pre++; }
            const size_t stop = (i + 1 < internal.size()) ? internal[i + 1] : order.size();
            for (size_t j = internal[i] + 1; j < stop; ++j) { out.leaves.push_back(order[j]); op.n_post++; }
            out.ops.push_back(op);
            out.n_gemm++;
        }
    }
};

// ------------------------------------------------------------------------------------------------------------------
// Reconstruction schedule: post-order with Sethi-Ullman ordering of internal children over a fixed number of
// shared-memory vector slots; vectors beyond the slots are spilled to an L2-resident scratch area.
// ------------------------------------------------------------------------------------------------------------------
struct Schedule {
    std::vector<Op> ops;
    int n_spill = 0;
};

class ScheduleBuilder {
public:
    ScheduleBuilder(const HostTree& t, int slots) : tree(t), n_slots(slots), owner(slots, -1) { compute_need(); }

    Schedule build()
    {
        int root = tree.n_nodes - 1;
        int vid = emit(root);
        make_resident(vid, -1);
        out.ops.push_back({OP_ROOT, where[vid], 0, root});
        return out;
    }

private:
    const HostTree& tree;
    int n_slots;
    std::vector<int> owner;                 // physical slot -> vector id
    std::vector<int> where;                 // vector id -> slot (>= 0) or -(spill index + 1)
    std::vector<int> birth;                 // vector id -> creation order (victim choice: oldest)
    std::vector<int> free_spill;
    std::vector<int> need;
    Schedule out;
    int clock = 0;

    void compute_need()
    {
        need.assign(tree.n_nodes, 0);
        for (int v = 0; v < tree.n_nodes; ++v) {
            if (tree.is_leaf(v)) continue;
            std::vector<int> ns;
            for (int e = tree.child_offset[v]; e < tree.child_offset[v + 1]; ++e) {
                int c = tree.child_list[e];
                if (!tree.is_leaf(c)) ns.push_back(need[c]);
            }
            std::sort(ns.rbegin(), ns.rend());
            int n = 1;
            for (size_t i = 0; i < ns.size(); ++i) n = std::max(n, ns[i] + (i > 0 ? 1 : 0));
            need[v] = n;
        }
    }

    int new_vector()
    {
        where.push_back(-1000000);
        birth.push_back(clock++);
        return (int)where.size() - 1;
    }

    int acquire(int pin_a, int pin_b)
    {
        for (int s = 0; s < n_slots; ++s)
            if (owner[s] < 0) return s;
        int victim = -1;
        for (int s = 0; s < n_slots; ++s) {
            int vid = owner[s];
            if (vid == pin_a || vid == pin_b) continue;
            if (victim < 0 || birth[vid] < birth[owner[victim]]) victim = s;
        }
        int idx;
        if (!free_spill.empty()) { idx = free_spill.back(); free_spill.pop_back(); }
        else idx = out.n_spill++;
        out.ops.push_back({OP_SPILL, victim, idx, 0});
        where[owner[victim]] = -(idx + 1);
        owner[victim] = -1;
        return victim;
    }

    void make_resident(int vid, int pin)
    {
        if (where[vid] >= 0) return;
        int idx = -where[vid] - 1;
        int s = acquire(vid, pin);
        out.ops.push_back({OP_FILL, s, idx, 0});
        free_spill.push_back(idx);
        where[vid] = s;
        owner[s] = vid;
    }

    void release(int vid)
    {
        owner[where[vid]] = -1;
        where[vid] = -1000000;
    }

    // Binary nodes: internal child with the larger need first, leaves last (a*b == b*a exactly, so
    // the product is bit-identical to the reference's child order).  Nodes with more than two
    // children keep Newick order, because the reference multiplies factors in that order
    // (src/gene_family_reconstructor.cpp:96-101) and a reassociated product could differ in the last bit.
    int emit(int v)
    {
        std::vector<int> order;
        for (int e = tree.child_offset[v]; e < tree.child_offset[v + 1]; ++e) order.push_back(tree.child_list[e]);
        if (order.size() <= 2)
            std::stable_sort(order.begin(), order.end(), [this](int a, int b) {
                const int na = tree.is_leaf(a) ? -1 : need[a], nb = tree.is_leaf(b) ? -1 : need[b];
                return na > nb;
            });
        int acc = -1;
        for (int c : order) {
            if (!tree.is_leaf(c)) {
                int vid = emit(c);
                make_resident(vid, acc);
                if (acc < 0) {
                    out.ops.push_back({OP_GEMM_SET, where[vid], where[vid], c});
                    acc = vid;
                }
                else {
                    make_resident(acc, vid);
                    out.ops.push_back({OP_GEMM_MUL, where[acc], where[vid], c});
                    release(vid);
                }
            }
            else if (acc < 0) {
                acc = new_vector();
                int s = acquire(-1, -1);
                where[acc] = s;
                owner[s] = acc;
                out.ops.push_back({OP_LEAF_SET, s, 0, c});
            }
            else {
                make_resident(acc, -1);
                out.ops.push_back({OP_LEAF_MUL, where[acc], 0, c});
            }
        }
        make_resident(acc, -1);
        out.ops.push_back({OP_RESCALE, where[acc], 0, v});
        return acc;
    }
};

}  // namespace cafe
