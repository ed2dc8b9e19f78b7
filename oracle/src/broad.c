/* TEST INFRASTRUCTURE (CPU oracle): AABB overlap, incremental AABB tree.
 * Follows distance3d/aabb_tree.py (cited per function). */
#include <stdlib.h>
#include <string.h>
#include "d3d_oracle.h"

/* aabb_tree.py:503-527 (closed intervals: touching boxes overlap) */
static inline int aabb_overlap(const double *a, const double *b) {
    return a[0] <= b[1] && a[1] >= b[0] && a[2] <= b[3] && a[3] >= b[2] && a[4] <= b[5] &&
           a[5] >= b[4];
}

/* aabb_tree.py:465-500 */
int64_t d3do_all_aabbs_overlap(const double *aabbs1, int64_t n1, const double *aabbs2,
                               int64_t n2, int32_t *out_pairs, int64_t cap) {
    int64_t count = 0;
    for (int64_t i = 0; i < n1; ++i)
        for (int64_t j = 0; j < n2; ++j)
            if (aabb_overlap(aabbs1 + 6 * i, aabbs2 + 6 * j)) {
                if (out_pairs && count < cap) {
                    out_pairs[2 * count] = (int32_t)i;
                    out_pairs[2 * count + 1] = (int32_t)j;
                }
                ++count;
            }
    return count;
}

enum { PARENT = 0, LEFT = 1, RIGHT = 2, TYPE = 3, T_LEAF = 1, T_BRANCH = 2, NONE = -1 };

/* aabb_tree.py:536-551 */
static void merge_aabb(const double *a, const double *b, double *o) {
    for (int k = 0; k < 3; ++k) {
        o[2 * k] = a[2 * k] < b[2 * k] ? a[2 * k] : b[2 * k];             /* min(a, b) */
        o[2 * k + 1] = a[2 * k + 1] > b[2 * k + 1] ? a[2 * k + 1] : b[2 * k + 1]; /* max(a, b) */
    }
}
static double merged_volume(const double *a, const double *b) {
    double m[6];
    merge_aabb(a, b, m);
    return (m[1] - m[0]) * (m[3] - m[2]) * (m[5] - m[4]);
}

/* aabb_tree.py:194-341 insert_aabbs / insert_leaf / fix_upward_tree */
int64_t d3do_tree_insert(int64_t root, int64_t *nodes, double *aabbs, int64_t *filled_len,
                         const int64_t *insert_order, int64_t n_insert) {
    int64_t filled;
    /* In the reference `filled_len` handed to insert_aabbs already counts the new
     * leaves (aabb_tree.py:57,93); branch nodes are appended after them. */
    filled = *filled_len;
    for (int64_t q = 0; q < n_insert; ++q) {
        int64_t leaf = insert_order[q];
        nodes[4 * leaf + TYPE] = T_LEAF;
        if (root == NONE) { root = leaf; continue; }
        int64_t t = root;
        while (nodes[4 * t + TYPE] == T_BRANCH) {
            int64_t l = nodes[4 * t + LEFT], r = nodes[4 * t + RIGHT];
            double cost_left = merged_volume(aabbs + 6 * leaf, aabbs + 6 * l);
            double cost_right = merged_volume(aabbs + 6 * leaf, aabbs + 6 * r);
            t = (cost_left < cost_right) ? l : r;
        }
        int64_t sibling = t;
        int64_t old_parent = nodes[4 * sibling + PARENT];
        int64_t np_ = filled++;
        nodes[4 * np_ + PARENT] = old_parent;
        nodes[4 * np_ + LEFT] = sibling;
        nodes[4 * np_ + RIGHT] = leaf;
        nodes[4 * np_ + TYPE] = T_BRANCH;
        merge_aabb(aabbs + 6 * leaf, aabbs + 6 * sibling, aabbs + 6 * np_);
        nodes[4 * leaf + PARENT] = np_;
        nodes[4 * sibling + PARENT] = np_;
        if (old_parent == NONE) root = np_;
        else if (nodes[4 * old_parent + LEFT] == sibling) nodes[4 * old_parent + LEFT] = np_;
        else nodes[4 * old_parent + RIGHT] = np_;
        int64_t u = nodes[4 * leaf + PARENT];
        while (u != NONE) {
            merge_aabb(aabbs + 6 * nodes[4 * u + LEFT], aabbs + 6 * nodes[4 * u + RIGHT],
                       aabbs + 6 * u);
            u = nodes[4 * u + PARENT];
        }
    }
    *filled_len = filled;
    return root;
}

typedef struct { int64_t *data; int64_t size, cap; } stack_t_;
static void push(stack_t_ *s, int64_t v) {
    if (s->size == s->cap) { s->cap = s->cap ? 2 * s->cap : 64; s->data = realloc(s->data, s->cap * sizeof(int64_t)); }
    s->data[s->size++] = v;
}

/* aabb_tree.py:381-403; returns number of overlapping leaves, calls emit for each */
static int64_t query_one(const double *test, int64_t root, const int64_t *nodes,
                         const double *aabbs, int break_at_first_leaf, stack_t_ *st,
                         int32_t *out_pairs, int64_t cap, int64_t base, int32_t tag) {
    int64_t count = 0;
    st->size = 0;
    if (root == NONE) return 0;
    push(st, root);
    while (st->size) {
        int64_t n = st->data[--st->size];
        if (aabb_overlap(aabbs + 6 * n, test)) {
            if (nodes[4 * n + TYPE] == T_LEAF) {
                if (out_pairs && base + count < cap) {
                    out_pairs[2 * (base + count)] = (int32_t)n;
                    out_pairs[2 * (base + count) + 1] = tag;
                }
                ++count;
                if (break_at_first_leaf) break;
            } else {
                push(st, nodes[4 * n + LEFT]);
                push(st, nodes[4 * n + RIGHT]);
            }
        }
    }
    return count;
}

int64_t d3do_tree_query(int64_t root, const int64_t *nodes, const double *aabbs,
                        const double *query, int64_t n_query, int32_t *out_pairs, int64_t cap,
                        int n_threads) {
    (void)n_threads; /* order of the emitted list is part of what tests compare: serial */
    stack_t_ st = {0, 0, 0};
    int64_t total = 0;
    for (int64_t q = 0; q < n_query; ++q)
        total += query_one(query + 6 * q, root, nodes, aabbs, 0, &st, out_pairs, cap, total,
                           (int32_t)q);
    free(st.data);
    return total;
}

/* aabb_tree.py:344-378 */
int64_t d3do_tree_vs_tree(int64_t root1, const int64_t *nodes1, const double *aabbs1,
                          int64_t root2, const int64_t *nodes2, const double *aabbs2,
                          int32_t *out_pairs, int64_t cap) {
    stack_t_ outer = {0, 0, 0}, inner = {0, 0, 0};
    int64_t total = 0;
    if (root2 == NONE) return 0;
    push(&outer, root2);
    while (outer.size) {
        int64_t n = outer.data[--outer.size];
        const double *box = aabbs2 + 6 * n;
        if (nodes2[4 * n + TYPE] == T_BRANCH &&
            query_one(box, root1, nodes1, aabbs1, 1, &inner, 0, 0, 0, 0) >= 1) {
            push(&outer, nodes2[4 * n + LEFT]);
            push(&outer, nodes2[4 * n + RIGHT]);
        } else if (nodes2[4 * n + TYPE] == T_LEAF) {
            total += query_one(box, root1, nodes1, aabbs1, 0, &inner, out_pairs, cap, total,
                               (int32_t)n);
        }
    }
    free(outer.data);
    free(inner.data);
    return total;
}
