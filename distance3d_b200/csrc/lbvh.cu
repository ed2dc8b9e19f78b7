// Broad phase on the GPU: linear BVH over fp64 AABBs + all-overlap queries.
//
// Replaces the reference's incremental AABB tree (distance3d/aabb_tree.py:194-341
// insert_aabbs / insert_leaf / fix_upward_tree) and its list-as-stack traversals
// (query_overlap :381-403, query_overlap_of_other_tree :344-378) as well as the
// brute-force all_aabbs_overlap (:465-500).  Only the overlap SET is contractual
// (the predicate aabb_overlap :503-527 is closed: touching boxes overlap); the
// tree shape is ours:
//
//   k_bounds_*      centroid bounds of the scene (two-stage min/max reduction)
//   k_morton        key = 30-bit Morton code of the centroid << 32 | object index
//   radix sort      hand-written LSD sort, 4 passes x 8 bits over the Morton bits:
//                   per-block digit histogram, scan, stable scatter that ranks with
//                   __match_any_sync (no CUB)
//   k_leaves        leaf records in sorted order
//   k_karras        Karras 2012 hierarchy: one thread per internal node, binary
//                   search on the common prefix of the (unique) keys; also emits the
//                   "rope" (next node in depth-first order when a subtree is skipped)
//   k_refit         bottom-up AABB union, second arrival proceeds (atomic flags)
//   k_overlap<F>    stackless traversal: node = overlap ? first child : rope.  Two passes:
//                   count hits per query, exclusive scan, then fill -- no atomics, exact
//                   count before anything is written, deterministic order
//   k_brute         brute force; its pairs are appended with a warp-aggregated atomic
//                   (ballot + one atomic per converged group)
//
// Build record = 64 bytes: exact fp64 lo[3], hi[3], left, right, rope, parent.  Internal node
// i in [0, n-1), leaf j (sorted position) at n-1+j; a leaf stores left = -(object index + 1).
//
// The queries read a second, compact copy of the tree.  Internal nodes: 32-byte records (two
// 128-bit loads) with the box rounded OUTWARD to fp32 plus left / rope; the six comparisons of
// a step are fp32 and the box is conservative, it can only add candidates.  Leaves: 64-byte
// records with the exact fp64 box of the object, its index and the rope, so a leaf is decided
// by the reference's closed-interval fp64 predicate in the same single fetch.  96 MB for 1 M
// boxes instead of 128 MB: the whole tree stays in the 126 MB L2.
#include "d3d_common.cuh"

namespace {

struct __align__(16) BvhNode {
    double lo[3];
    double hi[3];
    int left, right, rope, parent;
};
static_assert(sizeof(BvhNode) == 64, "node record must be 64 bytes");

struct __align__(16) TNode {  // traversal record of an internal node
    float lo[3];
    float hi[3];
    int left, rope;
};
static_assert(sizeof(TNode) == 32, "traversal record must be 32 bytes");
struct __align__(16) LeafRec {  // traversal record of a leaf (sorted position): the exact box
    double box[6];              // lo.x hi.x lo.y hi.y lo.z hi.z (the (3,2) layout of the input)
    int obj, rope;
    int pad[2];
};
static_assert(sizeof(LeafRec) == 64, "leaf record must be 64 bytes");

struct BvhHeader {  // first 256 bytes of the workspace
    int64_t n;
    int root;
    int pad;
    double bounds[6];
};

#define SORT_TILE 4096
#define SORT_THREADS 256
#define SORT_ITEMS (SORT_TILE / SORT_THREADS)

struct BvhLayout {
    BvhHeader *hdr;
    double *partials;             // [1024][6]
    unsigned long long *keys[2];  // [n] each
    unsigned *hist;               // [256][sort_blocks]
    unsigned *digit_totals;       // [256]
    int *flags;                   // [n-1]
    int *range_last;              // [2n-1] last sorted leaf covered by each node
    BvhNode *nodes;               // [2n-1]
    TNode *tnodes;                // [n-1] compact internal nodes for the queries
    LeafRec *leaves;              // [n] leaf records in sorted order
    int sort_blocks;
};

inline size_t au(size_t x) { return (x + 255) / 256 * 256; }

inline size_t bvh_ws_bytes(int64_t n) {
    size_t nn = (size_t)(n > 0 ? n : 1);
    size_t sort_blocks = (nn + SORT_TILE - 1) / SORT_TILE;
    return 256 + au(1024 * 6 * 8) + 1024 + 2 * au(nn * 8) + au(256 * sort_blocks * 4) + au(nn * 4) +
           au(2 * nn * 4) + au(2 * nn * sizeof(BvhNode)) + au(nn * sizeof(TNode)) + au(nn * sizeof(LeafRec));
}

inline BvhLayout bvh_carve(void *ws, int64_t n) {
    size_t nn = (size_t)(n > 0 ? n : 1);
    BvhLayout L;
    char *p = reinterpret_cast<char *>(ws);
    L.hdr = reinterpret_cast<BvhHeader *>(p); p += 256;
    L.partials = reinterpret_cast<double *>(p); p += au(1024 * 6 * 8);
    L.digit_totals = reinterpret_cast<unsigned *>(p); p += 1024;
    L.keys[0] = reinterpret_cast<unsigned long long *>(p); p += au(nn * 8);
    L.keys[1] = reinterpret_cast<unsigned long long *>(p); p += au(nn * 8);
    L.sort_blocks = (int)((nn + SORT_TILE - 1) / SORT_TILE);
    L.hist = reinterpret_cast<unsigned *>(p); p += au(256 * (size_t)L.sort_blocks * 4);
    L.flags = reinterpret_cast<int *>(p); p += au(nn * 4);
    L.range_last = reinterpret_cast<int *>(p); p += au(2 * nn * 4);
    L.nodes = reinterpret_cast<BvhNode *>(p); p += au(2 * nn * sizeof(BvhNode));
    L.tnodes = reinterpret_cast<TNode *>(p); p += au(nn * sizeof(TNode));
    L.leaves = reinterpret_cast<LeafRec *>(p);
    return L;
}

// ---------------------------------------------------------------- bounds
__global__ void k_bounds_partial(const double *__restrict__ aabb, int64_t n, double *partials) {
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const double2 *b = reinterpret_cast<const double2 *>(aabb + 6 * i);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            double2 v = __ldg(b + k);
            double c = 0.5 * (v.x + v.y);
            lo[k] = fmin(lo[k], c);
            hi[k] = fmax(hi[k], c);
        }
    }
    __shared__ double sh[256 / 32][6];
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 3; ++k)
        for (int off = 16; off > 0; off >>= 1) {
            lo[k] = fmin(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], off));
            hi[k] = fmax(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], off));
        }
    if (lane == 0)
        for (int k = 0; k < 3; ++k) { sh[wid][k] = lo[k]; sh[wid][3 + k] = hi[k]; }
    __syncthreads();
    if (threadIdx.x < 6) {
        double v = sh[0][threadIdx.x];
        for (int w = 1; w < blockDim.x / 32; ++w)
            v = threadIdx.x < 3 ? fmin(v, sh[w][threadIdx.x]) : fmax(v, sh[w][threadIdx.x]);
        partials[blockIdx.x * 6 + threadIdx.x] = v;
    }
}

__global__ void __launch_bounds__(256)
k_bounds_final(const double *partials, int n_partials, BvhHeader *hdr, int64_t n) {
    // 256 threads: component c = threadIdx.x % 6 of partial rows threadIdx.x / 6, +42, ...
    __shared__ double sh[252];
    int c = threadIdx.x % 6, r0 = threadIdx.x / 6;
    if (threadIdx.x < 252) {
        double v = c < 3 ? 1e300 : -1e300;
        for (int b = r0; b < n_partials; b += 42) {
            double x = partials[b * 6 + c];
            v = c < 3 ? fmin(v, x) : fmax(v, x);
        }
        sh[threadIdx.x] = v;
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        double v = sh[threadIdx.x];
        for (int r = 1; r < 42; ++r)
            v = threadIdx.x < 3 ? fmin(v, sh[r * 6 + threadIdx.x]) : fmax(v, sh[r * 6 + threadIdx.x]);
        hdr->bounds[threadIdx.x] = v;
    }
    if (threadIdx.x == 0) { hdr->n = n; hdr->root = 0; }
}

__device__ __forceinline__ unsigned expand_bits10(unsigned v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

__global__ void k_morton(const double *__restrict__ aabb, int64_t n, const BvhHeader *hdr,
                         unsigned long long *keys) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double2 *b = reinterpret_cast<const double2 *>(aabb + 6 * i);
    unsigned q[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double2 v = __ldg(b + k);
        double c = 0.5 * (v.x + v.y);
        double lo = hdr->bounds[k], ext = hdr->bounds[3 + k] - lo;
        double t = ext > 0.0 ? (c - lo) / ext * 1024.0 : 0.0;
        int g = (int)t;
        q[k] = (unsigned)min(max(g, 0), 1023);
    }
    unsigned code = (expand_bits10(q[0]) << 2) | (expand_bits10(q[1]) << 1) | expand_bits10(q[2]);
    keys[i] = ((unsigned long long)code << 32) | (unsigned long long)(unsigned)i;
}

// ---------------------------------------------------------------- radix sort
// Pass over digit bits [shift, shift+8).  hist is digit-major: hist[d * blocks + b].
__global__ void __launch_bounds__(SORT_THREADS)
k_sort_hist(const unsigned long long *__restrict__ keys, int64_t n, int shift, unsigned *hist,
            int blocks) {
    __shared__ unsigned sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    int64_t base = (int64_t)blockIdx.x * SORT_TILE;
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        int64_t k = base + i * SORT_THREADS + threadIdx.x;
        if (k < n) atomicAdd(&sh[(unsigned)(keys[k] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[threadIdx.x * blocks + blockIdx.x] = sh[threadIdx.x];
}

// Scan of the digit-major histogram hist[d * blocks + b] in two levels: one WARP per digit
// turns its row into an exclusive prefix over the blocks and records the digit total; the
// scatter kernel adds the exclusive prefix over the 256 digit totals (a block-local scan).
__global__ void __launch_bounds__(256) k_sort_scan(unsigned *hist, int blocks, unsigned *digit_totals) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= 256) return;
    unsigned *row = hist + (size_t)warp * blocks;
    unsigned carry = 0;
    for (int base = 0; base < blocks; base += 32) {
        int i = base + lane;
        unsigned v = i < blocks ? row[i] : 0u, x = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            unsigned y = __shfl_up_sync(0xffffffffu, x, off);
            if (lane >= off) x += y;
        }
        if (i < blocks) row[i] = carry + x - v;
        carry += __shfl_sync(0xffffffffu, x, 31);
    }
    if (lane == 0) digit_totals[warp] = carry;
}

// Stable scatter.  Warp w of the block owns the contiguous slice
// [w * 512, (w + 1) * 512) of the tile and walks it 32 keys at a time; equal
// digits inside one step are ranked with __match_any_sync.
__global__ void __launch_bounds__(SORT_THREADS)
k_sort_scatter(const unsigned long long *__restrict__ in, unsigned long long *__restrict__ out,
               int64_t n, int shift, const unsigned *__restrict__ hist, int blocks,
               const unsigned *__restrict__ digit_totals) {
    constexpr int WARPS = SORT_THREADS / 32;
    constexpr int PER_WARP = SORT_TILE / WARPS;
    constexpr int STEPS = PER_WARP / 32;
    __shared__ unsigned counts[WARPS][256];
    __shared__ unsigned digit_base[256];
    __shared__ unsigned warp_tot[WARPS];
    for (int i = threadIdx.x; i < WARPS * 256; i += SORT_THREADS) (&counts[0][0])[i] = 0;
    {   // exclusive scan of the 256 digit totals (thread d owns digit d)
        unsigned v = digit_totals[threadIdx.x], x = v;
        int ln = threadIdx.x & 31, wd = threadIdx.x >> 5;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            unsigned y = __shfl_up_sync(0xffffffffu, x, off);
            if (ln >= off) x += y;
        }
        if (ln == 31) warp_tot[wd] = x;
        __syncthreads();
        unsigned before = 0;
        for (int w2 = 0; w2 < wd; ++w2) before += warp_tot[w2];
        digit_base[threadIdx.x] = before + x - v;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1;
    int64_t base = (int64_t)blockIdx.x * SORT_TILE + (int64_t)wid * PER_WARP;
    unsigned long long key[STEPS];
    unsigned rank[STEPS];
#pragma unroll
    for (int s = 0; s < STEPS; ++s) {
        int64_t k = base + s * 32 + lane;
        bool valid = k < n;
        key[s] = valid ? in[k] : ~0ull;
        unsigned d = valid ? ((unsigned)(key[s] >> shift) & 255u) : 256u;
        unsigned peers = __match_any_sync(0xffffffffu, d);
        unsigned before = 0;
        if (valid) {
            int leader = __ffs(peers) - 1;
            if (lane == leader) {
                before = counts[wid][d];
                counts[wid][d] = before + __popc(peers);
            }
            before = __shfl_sync(peers, before, leader);
            rank[s] = before + __popc(peers & lt);
        }
        __syncwarp();
    }
    __syncthreads();
    // exclusive prefix over the warps of this block, per digit, plus the global offset
    {
        unsigned d = threadIdx.x;
        unsigned acc = digit_base[d] + hist[d * blocks + blockIdx.x];
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            unsigned c = counts[w][d];
            counts[w][d] = acc;
            acc += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < STEPS; ++s) {
        int64_t k = base + s * 32 + lane;
        if (k < n) {
            unsigned d = (unsigned)(key[s] >> shift) & 255u;
            out[counts[wid][d] + rank[s]] = key[s];
        }
    }
}

// ---------------------------------------------------------------- hierarchy
__device__ __forceinline__ int delta(const unsigned long long *keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    return __clzll(keys[i] ^ keys[j]);  // keys are unique (object index in the low bits)
}

// Leaf records (thread j < n) and the Karras hierarchy (thread i < n - 1) in one launch.  The
// leaf part does not touch `parent`: that field is written by the internal node that adopts it.
__device__ __forceinline__ void store_tbox(TNode *t, const double *lo, const double *hi) {
    float4 a = make_float4(__double2float_rd(lo[0]), __double2float_rd(lo[1]), __double2float_rd(lo[2]),
                           __double2float_ru(hi[0]));
    float2 b = make_float2(__double2float_ru(hi[1]), __double2float_ru(hi[2]));
    *reinterpret_cast<float4 *>(t) = a;
    *reinterpret_cast<float2 *>(&t->hi[1]) = b;
}

__global__ void k_hierarchy(const double *__restrict__ aabb, const unsigned long long *__restrict__ keys,
                            int n, BvhNode *nodes, TNode *tnodes, LeafRec *leaves, int *range_first,
                            int *range_last, int *flags) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    {
        int obj = (int)(unsigned)(keys[i] & 0xffffffffull);
        const double2 *b = reinterpret_cast<const double2 *>(aabb + 6 * (int64_t)obj);
        double2 x = __ldg(b), y = __ldg(b + 1), z = __ldg(b + 2);
        double2 *lb = reinterpret_cast<double2 *>(leaves + i);
        lb[0] = x; lb[1] = y; lb[2] = z;
        leaves[i].obj = obj;
        if (n == 1) leaves[0].rope = -1;
        BvhNode *nd = nodes + (n - 1 + i);
        nd->lo[0] = x.x; nd->lo[1] = y.x; nd->lo[2] = z.x;
        nd->hi[0] = x.y; nd->hi[1] = y.y; nd->hi[2] = z.y;
        nd->left = -(obj + 1);
        nd->right = -1;
        nd->rope = -1;
        if (n == 1) nd->parent = -1;
        range_last[n - 1 + i] = i;
    }
    if (i >= n - 1) return;
    flags[i] = 0;
    int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = delta(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    int gamma = i + s * d + min(d, 0);
    int first = min(i, j), last = max(i, j);
    int left = (first == gamma) ? (n - 1 + gamma) : gamma;
    int right = (last == gamma + 1) ? (n - 1 + gamma + 1) : (gamma + 1);
    nodes[i].left = left;
    nodes[i].right = right;
    tnodes[i].left = left;
    nodes[left].parent = i;
    nodes[right].parent = i;
    if (i == 0) nodes[0].parent = -1;
    range_first[i] = first;
    range_last[i] = last;
}

// rope(node covering [a, b]) = node that starts at b + 1: internal node b + 1 when its
// range grows to the right, else leaf b + 1; -1 behind the last leaf.
__global__ void k_ropes(const unsigned long long *__restrict__ keys, int n, BvhNode *nodes, TNode *tnodes,
                        LeafRec *leaves, const int *__restrict__ range_last) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= 2 * n - 1) return;
    int b = range_last[v];
    int rope;
    if (b + 1 >= n) rope = -1;
    else if (b + 1 >= n - 1) rope = n - 1 + (b + 1);
    else {
        int t = b + 1;
        int d = (delta(keys, n, t, t + 1) - delta(keys, n, t, t - 1)) >= 0 ? 1 : -1;
        rope = d > 0 ? t : n - 1 + t;
    }
    nodes[v].rope = rope;
    if (v < n - 1) tnodes[v].rope = rope;
    else leaves[v - (n - 1)].rope = rope;
}

// Bottom-up refit: every leaf thread climbs, the second thread to arrive at a node merges the
// two child boxes and goes on.  A node whose leaf range lies inside the 256 leaves of this
// block can only be reached by threads of this block: its arrival flag lives in shared memory
// and block-scope fences order the box stores (99.6 % of the nodes); only the nodes that span
// blocks pay for device-scope fences and global atomics.
__global__ void __launch_bounds__(256)
k_refit(int n, BvhNode *nodes, TNode *tnodes, int *flags, const int *__restrict__ range_first,
        const int *__restrict__ range_last) {
    __shared__ int sflag[256];
    const int b0 = blockIdx.x * 256;
    sflag[threadIdx.x] = 0;
    __syncthreads();
    int j = b0 + threadIdx.x;
    if (j >= n) return;
    int node = n - 1 + j;
    double lo[3], hi[3];  // box of the subtree this thread has finished (kept in registers)
    for (int k = 0; k < 3; ++k) { lo[k] = nodes[node].lo[k]; hi[k] = nodes[node].hi[k]; }
    int cur = nodes[node].parent;
    while (cur >= 0) {
        // my subtree's box is visible before I announce it; the first arrival stops
        if (__ldg(range_first + cur) >= b0 && __ldg(range_last + cur) < b0 + 256) {
            __threadfence_block();
            if (atomicAdd(&sflag[cur - b0], 1) == 0) return;
            __threadfence_block();
        } else {
            __threadfence();
            if (atomicAdd(&flags[cur], 1) == 0) return;
            __threadfence();
        }
        int l = nodes[cur].left, r = nodes[cur].right;
        const volatile BvhNode *sib = nodes + (l == node ? r : l);
        for (int k = 0; k < 3; ++k) {
            lo[k] = fmin(lo[k], sib->lo[k]);
            hi[k] = fmax(hi[k], sib->hi[k]);
            nodes[cur].lo[k] = lo[k];
            nodes[cur].hi[k] = hi[k];
        }
        store_tbox(tnodes + cur, lo, hi);
        node = cur;
        cur = nodes[cur].parent;
    }
}

// ---------------------------------------------------------------- traversal
__device__ __forceinline__ void append_pair(int a, int b, int32_t *out_pairs, int64_t cap,
                                            unsigned long long *count) {
    unsigned m = __activemask();
    int lane = threadIdx.x & 31;
    int leader = __ffs(m) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(count, (unsigned long long)__popc(m));
    base = __shfl_sync(m, base, leader);
    unsigned long long pos = base + __popc(m & ((1u << lane) - 1));
    if ((int64_t)pos < cap) reinterpret_cast<int2 *>(out_pairs)[pos] = make_int2(a, b);
}

// One traversal step on the compact tree.  q*: exact fp64 query box ((lo, hi) per axis),
// f*: the same box rounded outward to fp32.
struct QueryBox {
    double2 x, y, z;
    float lox, loy, loz, hix, hiy, hiz;
    __device__ __forceinline__ void set(double2 qx, double2 qy, double2 qz) {
        x = qx; y = qy; z = qz;
        lox = __double2float_rd(qx.x); hix = __double2float_ru(qx.y);
        loy = __double2float_rd(qy.x); hiy = __double2float_ru(qy.y);
        loz = __double2float_rd(qz.x); hiz = __double2float_ru(qz.y);
    }
    __device__ __forceinline__ void set_empty() {  // never overlaps anything
        x = y = z = make_double2(1e308, -1e308);
        lox = loy = loz = 3.0e38f; hix = hiy = hiz = -3.0e38f;
    }
};
struct Step {
    bool ov;      // internal node: the fp32 boxes overlap (conservative); leaf: the exact boxes overlap
    int left;     // >= 0: first child; < 0: leaf of object -left - 1
    int rope;
};
struct Tree {
    const TNode *__restrict__ tnodes;
    const LeafRec *__restrict__ leaves;
    int leaf0;  // node index of the first leaf (n - 1)
};
// Whether `node` is a leaf is known from its index, so a leaf costs ONE fetch: its 64-byte
// record holds the exact box (aabb_tree.py:520-527, closed intervals), the object and the rope.
__device__ __forceinline__ Step visit(const Tree &T, int node, const QueryBox &q) {
    Step s;
    if (node >= T.leaf0) {
        const double2 *lb = reinterpret_cast<const double2 *>(T.leaves + (node - T.leaf0));
        double2 x = __ldg(lb), y = __ldg(lb + 1), z = __ldg(lb + 2);
        int2 link = __ldg(reinterpret_cast<const int2 *>(lb + 3));
        s.ov = x.x <= q.x.y && x.y >= q.x.x && y.x <= q.y.y && y.y >= q.y.x && z.x <= q.z.y && z.y >= q.z.x;
        s.left = -link.x - 1;
        s.rope = link.y;
    } else {
        const float4 *p = reinterpret_cast<const float4 *>(T.tnodes + node);
        float4 a = __ldg(p), b = __ldg(p + 1);  // lo.x lo.y lo.z hi.x | hi.y hi.z left rope
        s.ov = a.x <= q.hix && a.w >= q.lox && a.y <= q.hiy && b.x >= q.loy && a.z <= q.hiz && b.y >= q.loz;
        s.left = __float_as_int(b.z);
        s.rope = __float_as_int(b.w);
    }
    return s;
}

static inline Tree tree_of(const BvhLayout &L, int64_t n) {
    Tree T;
    T.tnodes = L.tnodes; T.leaves = L.leaves; T.leaf0 = (int)n - 1;
    return T;
}

// Traversal of one query box.  FILL = false: count the overlapping leaves;
// FILL = true: write (object, query) pairs to out[offset ..).  Two passes instead of an
// atomic append: no atomics, an exact count before anything is written, and a deterministic
// output order (queries in the given order, leaves in depth-first order).
template <bool FILL>
__global__ void __launch_bounds__(128)
k_overlap(Tree T, const BvhHeader *hdr,
          const double *__restrict__ query, const int32_t *__restrict__ order, int64_t n_query,
          unsigned *__restrict__ counts, const unsigned long long *__restrict__ offsets,
          int32_t *out_pairs, int64_t cap, unsigned long long *visits) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n_query) return;
    int qi = order ? order[t] : (int)t;
    const double2 *qb = reinterpret_cast<const double2 *>(query + 6 * (int64_t)qi);
    QueryBox q;
    q.set(__ldg(qb), __ldg(qb + 1), __ldg(qb + 2));
    int node = hdr->n > 0 ? hdr->root : -1;
    unsigned n_hits = 0, n_visited = 0;
    unsigned long long pos = FILL ? offsets[t] : 0ull;
    while (node >= 0) {
        Step s = visit(T, node, q);
        if (s.ov && s.left < 0) {
            if (FILL) {
                if ((int64_t)pos < cap) reinterpret_cast<int2 *>(out_pairs)[pos] = make_int2(-s.left - 1, qi);
                ++pos;
            } else {
                ++n_hits;
            }
        }
        node = (s.ov && s.left >= 0) ? s.left : s.rope;
        ++n_visited;
    }
    if (!FILL) counts[t] = n_hits;
    if (!FILL && visits) {  // measurement only: node records fetched (roofline traffic term)
        unsigned total = n_visited;
        unsigned m = __activemask();
        if (m == 0xffffffffu) {
            for (int off = 16; off > 0; off >>= 1) total += __shfl_xor_sync(m, total, off);
            if ((threadIdx.x & 31) == 0) atomicAdd(visits, (unsigned long long)total);
        } else {
            atomicAdd(visits, (unsigned long long)n_visited);
        }
    }
}

// Packet traversal: the 32 queries of a warp (neighbours in Morton order) walk the tree
// TOGETHER.  The node is fetched once per warp (uniform address, one wavefront) and every
// lane tests its own box; the warp descends when ANY lane overlaps.  The union of the
// nodes seen by 32 coherent queries is a small multiple of what one query sees, so the
// L2 traffic drops by an order of magnitude on dense scenes.  Same results as k_overlap.
template <bool FILL>
__global__ void __launch_bounds__(128)
k_overlap_packet(Tree T, const BvhHeader *hdr,
                 const double *__restrict__ query, const int32_t *__restrict__ order, int64_t n_query,
                 unsigned *__restrict__ counts, const unsigned long long *__restrict__ offsets,
                 int32_t *out_pairs, int64_t cap, unsigned long long *visits) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    bool valid = t < n_query;
    int qi = valid ? (order ? order[t] : (int)t) : 0;
    QueryBox q;
    q.set_empty();
    if (valid) {
        const double2 *qb = reinterpret_cast<const double2 *>(query + 6 * (int64_t)qi);
        q.set(__ldg(qb), __ldg(qb + 1), __ldg(qb + 2));
    }
    int node = hdr->n > 0 ? hdr->root : -1;  // warp-uniform
    unsigned n_hits = 0, n_visited = 0;
    unsigned long long pos = (FILL && valid) ? offsets[t] : 0ull;
    while (node >= 0) {
        Step s = visit(T, node, q);
        if (s.left < 0) {  // leaf (uniform branch)
            if (s.ov) {
                if (FILL) {
                    if ((int64_t)pos < cap) reinterpret_cast<int2 *>(out_pairs)[pos] = make_int2(-s.left - 1, qi);
                    ++pos;
                } else {
                    ++n_hits;
                }
            }
            node = s.rope;
        } else {
            node = __any_sync(0xffffffffu, s.ov) ? s.left : s.rope;
        }
        ++n_visited;
    }
    if (!FILL && valid) counts[t] = n_hits;
    if (!FILL && visits && (threadIdx.x & 31) == 0) atomicAdd(visits, (unsigned long long)n_visited);
}

// Single-pass query: hits are staged per warp in shared memory and flushed with ONE atomic
// reservation per APPEND_STAGE pairs and coalesced 8-byte stores (the first version of this
// file reserved per leaf hit and waited on that atomic 80 % of the time).  The pair order
// depends on the order of the reservations; callers that need a reproducible order use the
// count / fill passes.  *cursor ends up as the exact number of pairs even when `cap` is too
// small (the excess is dropped).
//
// PW = packet width: groups of PW neighbouring queries (Morton order) walk the tree together
// and descend when ANY query of the group overlaps.  PW = 1 is the independent traversal,
// PW = 32 the full-warp packet; in between, the union of the nodes a group must see shrinks
// faster than the number of groups per warp grows (dense capsule set: 8-wide groups need
// ~2x fewer warp steps than 32-wide ones), while the PW lanes of a group still fetch one node.
//
// SELF = true: the queries are the tree's own leaves (sorted positions) and every unordered pair
// is wanted ONCE, without (i, i).  The leaves are dealt to `n_parts` parts (one per GPU) in
// blocks of 128 = one CTA, round-robin: CTA b of part p walks leaves (b * n_parts + p) * 128 ...
// (a contiguous split would give the part with the early leaves most of the pairs, because a
// pair belongs to its earlier leaf).  The depth-first
// order of the tree is the sorted leaf order, so query s only needs the part of the walk that
// lies to the right of leaf s: its group starts AT the leaf record of the group's first query
// and follows the ropes from there (a rope always leads to the subtree that begins behind the
// current one).  Half the nodes, half the output; pairs are written as (min, max) of the two
// object indices.
#ifndef APPEND_STAGE
#define APPEND_STAGE 512  // pairs per warp; dense / sparse ms at 128, 256, 512, 1024: 14.6 14.3 14.2 20.2 / 0.50 0.50 0.47 0.66
#endif
#ifndef D3D_BVH_PACKET_WIDTH
#define D3D_BVH_PACKET_WIDTH 8  // width used for packet = 1; 1 M capsules, ms dense / sparse at
// widths 1, 2, 4, 8, 16, 32: 34.8 23.0 16.7 14.0 15.5 14.9 / 0.61 0.54 0.48 0.47 0.58 0.65
#endif
template <int PW, bool SELF>
__global__ void __launch_bounds__(128)
k_overlap_append(Tree T, const BvhHeader *hdr,
                 const double *__restrict__ query, const int32_t *__restrict__ order, int part, int n_parts,
                 int64_t n_query, int32_t *out_pairs, int64_t cap, unsigned long long *cursor,
                 unsigned long long *visits) {
    __shared__ int2 stage_all[4][APPEND_STAGE];
    int2 *stage = stage_all[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu, lt = (1u << lane) - 1;
    const unsigned group_mask = (PW == 32) ? FULL : (((1u << PW) - 1u) << (lane & ~(PW - 1)));
    // SELF: t = sorted position of this thread's leaf, n_query = number of leaves of the tree
    int64_t t = (SELF ? (int64_t)blockIdx.x * n_parts + part : (int64_t)blockIdx.x) * (int64_t)blockDim.x + threadIdx.x;
    bool valid = t < n_query;
    int qi = 0;
    int my_leaf = 0x7fffffff;  // SELF: node index of this query's own leaf record
    QueryBox q;
    q.set_empty();
    if (valid) {
        const double2 *qb;
        if (SELF) {
            my_leaf = T.leaf0 + (int)t;
            qi = __ldg(&T.leaves[t].obj);
            qb = reinterpret_cast<const double2 *>(T.leaves + t);
        } else {
            qi = order ? order[t] : (int)t;
            qb = reinterpret_cast<const double2 *>(query + 6 * (int64_t)qi);
        }
        q.set(__ldg(qb), __ldg(qb + 1), __ldg(qb + 2));
    }
    // group-uniform; a group with no valid query at all never starts
    bool group_valid = PW == 1 ? valid : (__ballot_sync(FULL, valid) & group_mask) != 0;
    int node = -1;
    if (hdr->n > 0 && group_valid) {
        if (SELF) node = T.leaf0 + (int)(t & ~(int64_t)(PW - 1));  // the group's first leaf (always valid)
        else node = hdr->root;
    }
    int staged = 0;  // warp-uniform
    unsigned n_visited = 0;
    while (__any_sync(FULL, node >= 0)) {
        bool ov = false, hit = false;
        int left = 0, rope = -1;
        if (node >= 0) {
            Step s = visit(T, node, q);
            ov = s.ov; left = s.left; rope = s.rope;
            ++n_visited;
        }
        bool any = PW == 1 ? ov : (__ballot_sync(FULL, ov) & group_mask) != 0;
        int leaf = -left - 1;
        if (node >= 0) {
            hit = ov && left < 0 && (!SELF || node > my_leaf);
            node = (any && left >= 0) ? left : rope;
        }
        unsigned m = __ballot_sync(FULL, hit);
        if (m) {
            if (hit) stage[staged + __popc(m & lt)] = SELF ? make_int2(min(leaf, qi), max(leaf, qi)) : make_int2(leaf, qi);
            staged += __popc(m);
        }
        if (staged > APPEND_STAGE - 32) {
            __syncwarp();
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(cursor, (unsigned long long)staged);
            base = __shfl_sync(FULL, base, 0);
            for (int i = lane; i < staged; i += 32)
                if ((int64_t)(base + i) < cap) reinterpret_cast<int2 *>(out_pairs)[base + i] = stage[i];
            staged = 0;
            __syncwarp();
        }
    }
    if (staged) {
        __syncwarp();
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(cursor, (unsigned long long)staged);
        base = __shfl_sync(FULL, base, 0);
        for (int i = lane; i < staged; i += 32)
            if ((int64_t)(base + i) < cap) reinterpret_cast<int2 *>(out_pairs)[base + i] = stage[i];
    }
    if (visits) {  // node records fetched: one per group and step
        unsigned total = (lane & (PW - 1)) == 0 ? n_visited : 0u;
        for (int off = 16; off > 0; off >>= 1) total += __shfl_xor_sync(FULL, total, off);
        if (lane == 0) atomicAdd(visits, (unsigned long long)total);
    }
}

// ---- exclusive scan counts[u32] -> offsets[u64] (three small kernels) ----------
#define SCAN_TILE 4096  // 1024 threads x 4 items
__global__ void __launch_bounds__(1024)
k_scan_tiles(const unsigned *__restrict__ counts, int64_t n, unsigned long long *offsets,
             unsigned long long *tile_sums) {
    __shared__ unsigned long long warp_sums[32];
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + threadIdx.x * 4;
    unsigned long long v[4], sum = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[i] = base + i < n ? counts[base + i] : 0u; sum += v[i]; }
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned long long x = sum;
    for (int off = 1; off < 32; off <<= 1) {
        unsigned long long y = __shfl_up_sync(0xffffffffu, x, off);
        if (lane >= off) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
        unsigned long long s = warp_sums[lane];
        for (int off = 1; off < 32; off <<= 1) {
            unsigned long long y = __shfl_up_sync(0xffffffffu, s, off);
            if (lane >= off) s += y;
        }
        warp_sums[lane] = s;
    }
    __syncthreads();
    unsigned long long excl = (wid ? warp_sums[wid - 1] : 0ull) + x - sum;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (base + i < n) offsets[base + i] = excl;
        excl += v[i];
    }
    if (threadIdx.x == 1023) tile_sums[blockIdx.x] = excl;
}

__global__ void __launch_bounds__(1024)
k_scan_top(unsigned long long *tile_sums, int n_tiles, unsigned long long *total) {
    // single block, serial over chunks of 1024 tiles
    __shared__ unsigned long long warp_sums[32];
    __shared__ unsigned long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < n_tiles; base += 1024) {
        int i = base + threadIdx.x;
        unsigned long long v = i < n_tiles ? tile_sums[i] : 0ull, x = v;
        for (int off = 1; off < 32; off <<= 1) {
            unsigned long long y = __shfl_up_sync(0xffffffffu, x, off);
            if (lane >= off) x += y;
        }
        if (lane == 31) warp_sums[wid] = x;
        __syncthreads();
        if (wid == 0) {
            unsigned long long s = warp_sums[lane];
            for (int off = 1; off < 32; off <<= 1) {
                unsigned long long y = __shfl_up_sync(0xffffffffu, s, off);
                if (lane >= off) s += y;
            }
            warp_sums[lane] = s;
        }
        __syncthreads();
        unsigned long long prefix = carry + (wid ? warp_sums[wid - 1] : 0ull) + x - v;
        if (i < n_tiles) tile_sums[i] = prefix;
        __syncthreads();
        if (threadIdx.x == 1023) carry = prefix + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void k_scan_add(unsigned long long *offsets, int64_t n, const unsigned long long *tile_sums) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) offsets[i] += tile_sums[i / SCAN_TILE];
}

struct QueryLayout {
    unsigned *counts;
    unsigned long long *offsets;
    unsigned long long *tile_sums;
    int n_tiles;
};

inline size_t query_ws_bytes(int64_t nq) {
    size_t q = (size_t)(nq > 0 ? nq : 1);
    return au(q * 4) + au(q * 8) + au(((q + SCAN_TILE - 1) / SCAN_TILE) * 8);
}

inline QueryLayout query_carve(void *ws, int64_t nq) {
    size_t q = (size_t)(nq > 0 ? nq : 1);
    QueryLayout L;
    char *p = reinterpret_cast<char *>(ws);
    L.counts = reinterpret_cast<unsigned *>(p); p += au(q * 4);
    L.offsets = reinterpret_cast<unsigned long long *>(p); p += au(q * 8);
    L.tile_sums = reinterpret_cast<unsigned long long *>(p);
    L.n_tiles = (int)((q + SCAN_TILE - 1) / SCAN_TILE);
    return L;
}

// aabb_tree.py:465-500 all_aabbs_overlap: every (i, j) with overlapping boxes
__global__ void __launch_bounds__(256)
k_brute(const double *__restrict__ a1, int64_t n1, const double *__restrict__ a2, int64_t n2,
        int32_t *out_pairs, int64_t cap, unsigned long long *count) {
    __shared__ double tile[256][6];
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    double b[6];
    if (i < n1)
        for (int k = 0; k < 6; ++k) b[k] = a1[6 * i + k];
    for (int64_t j0 = (int64_t)blockIdx.y * 256; j0 < n2; j0 += (int64_t)gridDim.y * 256) {
        __syncthreads();
        int64_t j = j0 + threadIdx.x;
        if (j < n2)
            for (int k = 0; k < 6; ++k) tile[threadIdx.x][k] = a2[6 * j + k];
        __syncthreads();
        int lim = (int)d3d_min64(256, n2 - j0);
        if (i < n1)
            for (int jj = 0; jj < lim; ++jj) {
                const double *o = tile[jj];
                if (b[0] <= o[1] && b[1] >= o[0] && b[2] <= o[3] && b[3] >= o[2] && b[4] <= o[5] &&
                    b[5] >= o[4])
                    append_pair((int)i, (int)(j0 + jj), out_pairs, cap, count);
            }
    }
}

__global__ void k_leaf_order(const unsigned long long *__restrict__ keys, int64_t n, int32_t *out) {
    int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j < n) out[j] = (int32_t)(unsigned)(keys[j] & 0xffffffffull);
}

__global__ void k_root_aabb(const BvhNode *nodes, const BvhHeader *hdr, double *out) {
    if (threadIdx.x < 3) {
        out[2 * threadIdx.x] = nodes[hdr->root].lo[threadIdx.x];
        out[2 * threadIdx.x + 1] = nodes[hdr->root].hi[threadIdx.x];
    }
}

}  // namespace

extern "C" {

size_t d3d_bvh_workspace_bytes(int64_t n) { return bvh_ws_bytes(n); }

int d3d_bvh_build(const double *aabb, int64_t n, void *workspace, size_t ws_bytes, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!workspace) return d3d_set_error("d3d_bvh_build: null workspace");
    if (n < 0 || n > 0x3fffffff) return d3d_set_error("d3d_bvh_build: n out of range");
    if (ws_bytes < bvh_ws_bytes(n)) return d3d_set_error("d3d_bvh_build: workspace too small");
    BvhLayout L = bvh_carve(workspace, n);
    if (n == 0) {
        D3D_CUDA_CHECK(cudaMemsetAsync(L.hdr, 0, 256, stream));
        return 0;
    }
    if (!aabb) return d3d_set_error("d3d_bvh_build: null aabb");
    int pb = (int)d3d_min64((n + 255) / 256, 1024);
    k_bounds_partial<<<pb, 256, 0, stream>>>(aabb, n, L.partials);
    k_bounds_final<<<1, 256, 0, stream>>>(L.partials, pb, L.hdr, n);
    unsigned nb = (unsigned)((n + 255) / 256);
    k_morton<<<nb, 256, 0, stream>>>(aabb, n, L.hdr, L.keys[0]);
    int cur = 0;
    for (int shift = 32; shift < 64; shift += 8) {
        k_sort_hist<<<L.sort_blocks, SORT_THREADS, 0, stream>>>(L.keys[cur], n, shift, L.hist, L.sort_blocks);
        k_sort_scan<<<32, 256, 0, stream>>>(L.hist, L.sort_blocks, L.digit_totals);
        k_sort_scatter<<<L.sort_blocks, SORT_THREADS, 0, stream>>>(L.keys[cur], L.keys[cur ^ 1], n, shift,
                                                                  L.hist, L.sort_blocks, L.digit_totals);
        cur ^= 1;
    }
    // 4 passes: sorted keys are back in keys[0]
    // the second key buffer is free after the sort: it holds the first leaf of every internal node
    int *range_first = reinterpret_cast<int *>(L.keys[1]);
    k_hierarchy<<<nb, 256, 0, stream>>>(aabb, L.keys[0], (int)n, L.nodes, L.tnodes, L.leaves, range_first,
                                        L.range_last, L.flags);
    if (n > 1) {
        k_ropes<<<(unsigned)((2 * n - 1 + 255) / 256), 256, 0, stream>>>(L.keys[0], (int)n, L.nodes, L.tnodes,
                                                                       L.leaves, L.range_last);
        k_refit<<<nb, 256, 0, stream>>>((int)n, L.nodes, L.tnodes, L.flags, range_first, L.range_last);
    }
    D3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

size_t d3d_bvh_query_workspace_bytes(int64_t n_query) { return query_ws_bytes(n_query); }

/* pass 1: per-query hit counts + exclusive scan; *out_count = total number of pairs */
int d3d_bvh_overlap_count(const void *workspace, int64_t n, const double *query, const int32_t *order,
                          int64_t n_query, int packet, unsigned long long *out_count,
                          unsigned long long *out_visits, void *query_ws, size_t query_ws_size,
                          void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!workspace || !out_count) return d3d_set_error("d3d_bvh_overlap_count: null argument");
    D3D_CUDA_CHECK(cudaMemsetAsync(out_count, 0, sizeof(unsigned long long), stream));
    if (out_visits) D3D_CUDA_CHECK(cudaMemsetAsync(out_visits, 0, sizeof(unsigned long long), stream));
    if (n_query == 0 || n == 0) return 0;
    if (!query || !query_ws) return d3d_set_error("d3d_bvh_overlap_count: null query / workspace");
    if (query_ws_size < query_ws_bytes(n_query)) return d3d_set_error("d3d_bvh_overlap_count: query workspace too small");
    BvhLayout L = bvh_carve(const_cast<void *>(workspace), n);
    QueryLayout Q = query_carve(query_ws, n_query);
    if (packet)
        k_overlap_packet<false><<<(unsigned)((n_query + 127) / 128), 128, 0, stream>>>(
            tree_of(L, n), L.hdr, query, order, n_query, Q.counts, nullptr, nullptr, 0, out_visits);
    else
        k_overlap<false><<<(unsigned)((n_query + 127) / 128), 128, 0, stream>>>(
            tree_of(L, n), L.hdr, query, order, n_query, Q.counts, nullptr, nullptr, 0, out_visits);
    k_scan_tiles<<<Q.n_tiles, 1024, 0, stream>>>(Q.counts, n_query, Q.offsets, Q.tile_sums);
    k_scan_top<<<1, 1024, 0, stream>>>(Q.tile_sums, Q.n_tiles, out_count);
    k_scan_add<<<(unsigned)((n_query + 255) / 256), 256, 0, stream>>>(Q.offsets, n_query, Q.tile_sums);
    D3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

/* pass 2: writes the pairs (positions >= cap are dropped); needs the query workspace of pass 1 */
int d3d_bvh_overlap_fill(const void *workspace, int64_t n, const double *query, const int32_t *order,
                         int64_t n_query, int packet, int32_t *out_pairs, int64_t cap,
                         const void *query_ws, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_query == 0 || n == 0 || cap == 0) return 0;
    if (!workspace || !query || !out_pairs || !query_ws) return d3d_set_error("d3d_bvh_overlap_fill: null argument");
    BvhLayout L = bvh_carve(const_cast<void *>(workspace), n);
    QueryLayout Q = query_carve(const_cast<void *>(query_ws), n_query);
    if (packet)
        k_overlap_packet<true><<<(unsigned)((n_query + 127) / 128), 128, 0, stream>>>(
            tree_of(L, n), L.hdr, query, order, n_query, nullptr, Q.offsets, out_pairs, cap, nullptr);
    else
        k_overlap<true><<<(unsigned)((n_query + 127) / 128), 128, 0, stream>>>(
            tree_of(L, n), L.hdr, query, order, n_query, nullptr, Q.offsets, out_pairs, cap, nullptr);
    D3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

static int launch_append(const BvhLayout &L, int64_t n, bool self, const double *query, const int32_t *order,
                         int part, int n_parts, int64_t n_query, int packet, int32_t *out_pairs, int64_t cap,
                         unsigned long long *out_count, unsigned long long *out_visits, cudaStream_t stream) {
    int64_t all_blocks = (n_query + 127) / 128;
    unsigned blocks = (unsigned)(self ? (all_blocks - part + n_parts - 1) / n_parts : all_blocks);
    if (blocks == 0) return 0;
#define D3D_APPEND(PW)                                                                                  \
    do {                                                                                                \
        if (self)                                                                                       \
            k_overlap_append<PW, true><<<blocks, 128, 0, stream>>>(tree_of(L, n), L.hdr, query, order, \
                                                                   part, n_parts, n_query, out_pairs, cap, \
                                                                   out_count, out_visits);             \
        else                                                                                            \
            k_overlap_append<PW, false><<<blocks, 128, 0, stream>>>(tree_of(L, n), L.hdr, query, order, \
                                                                    part, n_parts, n_query, out_pairs, cap, \
                                                                    out_count, out_visits);            \
    } while (0)
    switch (packet) {  // 0 / 1 = per thread / default packet; 2, 4, 8, 16, 32 = explicit width
    case 0: D3D_APPEND(1); break;
    case 1: D3D_APPEND(D3D_BVH_PACKET_WIDTH); break;
    case 2: D3D_APPEND(2); break;
    case 4: D3D_APPEND(4); break;
    case 8: D3D_APPEND(8); break;
    case 16: D3D_APPEND(16); break;
    case 32: D3D_APPEND(32); break;
    default: return d3d_set_error("d3d_bvh_overlap: packet must be 0, 1, 2, 4, 8, 16 or 32");
    }
#undef D3D_APPEND
    D3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

/* single pass: traverse and append (unordered output, exact *out_count); query_ws is unused and
 * kept in the signature for the two-pass variant below */
int d3d_bvh_overlap(const void *workspace, int64_t n, const double *query, const int32_t *order,
                    int64_t n_query, int packet, int32_t *out_pairs, int64_t cap,
                    unsigned long long *out_count, unsigned long long *out_visits, void *query_ws,
                    size_t query_ws_size, void *stream_) {
    (void)query_ws; (void)query_ws_size;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!workspace || !out_count) return d3d_set_error("d3d_bvh_overlap: null argument");
    D3D_CUDA_CHECK(cudaMemsetAsync(out_count, 0, sizeof(unsigned long long), stream));
    if (out_visits) D3D_CUDA_CHECK(cudaMemsetAsync(out_visits, 0, sizeof(unsigned long long), stream));
    if (n_query == 0 || n == 0) return 0;
    if (!query || (cap > 0 && !out_pairs)) return d3d_set_error("d3d_bvh_overlap: null query / output");
    BvhLayout L = bvh_carve(const_cast<void *>(workspace), n);
    return launch_append(L, n, false, query, order, 0, 1, n_query, packet, out_pairs, cap, out_count, out_visits, stream);
}

/* the tree against its own leaves: every unordered pair once, as (smaller, larger) object index,
 * no (i, i).  Part `part` of `n_parts` (one per GPU; 0 of 1 = everything) takes the 128-leaf
 * blocks part, part + n_parts, ... of the sorted order. */
int d3d_bvh_overlap_self(const void *workspace, int64_t n, int part, int n_parts, int packet,
                         int32_t *out_pairs, int64_t cap, unsigned long long *out_count,
                         unsigned long long *out_visits, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!workspace || !out_count) return d3d_set_error("d3d_bvh_overlap_self: null argument");
    D3D_CUDA_CHECK(cudaMemsetAsync(out_count, 0, sizeof(unsigned long long), stream));
    if (out_visits) D3D_CUDA_CHECK(cudaMemsetAsync(out_visits, 0, sizeof(unsigned long long), stream));
    if (n == 0) return 0;
    if (n_parts < 1 || part < 0 || part >= n_parts) return d3d_set_error("d3d_bvh_overlap_self: part out of range");
    if (cap > 0 && !out_pairs) return d3d_set_error("d3d_bvh_overlap_self: null output");
    BvhLayout L = bvh_carve(const_cast<void *>(workspace), n);
    return launch_append(L, n, true, nullptr, nullptr, part, n_parts, n, packet, out_pairs, cap, out_count,
                         out_visits, stream);
}

/* count + fill in one call: reproducible pair order (queries in the given order, leaves in
 * depth-first order) at the price of a second traversal */
int d3d_bvh_overlap_ordered(const void *workspace, int64_t n, const double *query, const int32_t *order,
                            int64_t n_query, int packet, int32_t *out_pairs, int64_t cap,
                            unsigned long long *out_count, unsigned long long *out_visits,
                            void *query_ws, size_t query_ws_size, void *stream_) {
    int rc = d3d_bvh_overlap_count(workspace, n, query, order, n_query, packet, out_count, out_visits,
                                   query_ws, query_ws_size, stream_);
    if (rc) return rc;
    return d3d_bvh_overlap_fill(workspace, n, query, order, n_query, packet, out_pairs, cap, query_ws,
                                stream_);
}

/* sorted object order of the tree (Morton order): out[j] = object index of leaf j */
int d3d_bvh_leaf_order(const void *workspace, int64_t n, int32_t *out, void *stream_) {
    if (n == 0) return 0;
    if (!workspace || !out) return d3d_set_error("d3d_bvh_leaf_order: null argument");
    BvhLayout L = bvh_carve(const_cast<void *>(workspace), n);
    k_leaf_order<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(L.keys[0], n, out);
    D3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

/* root AABB (aabb_tree.py:183-191 get_root_aabb): out[3,2] */
int d3d_bvh_root_aabb(const void *workspace, int64_t n, double *out, void *stream_) {
    if (n == 0) return d3d_set_error("d3d_bvh_root_aabb: empty tree");
    if (!workspace || !out) return d3d_set_error("d3d_bvh_root_aabb: null argument");
    BvhLayout L = bvh_carve(const_cast<void *>(workspace), n);
    k_root_aabb<<<1, 32, 0, (cudaStream_t)stream_>>>(L.nodes, L.hdr, out);
    D3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int d3d_aabb_overlap_brute(const double *aabb1, int64_t n1, const double *aabb2, int64_t n2,
                           int32_t *out_pairs, int64_t cap, unsigned long long *out_count,
                           void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!out_count) return d3d_set_error("d3d_aabb_overlap_brute: null argument");
    D3D_CUDA_CHECK(cudaMemsetAsync(out_count, 0, sizeof(unsigned long long), stream));
    if (n1 == 0 || n2 == 0) return 0;
    dim3 grid((unsigned)((n1 + 255) / 256), (unsigned)d3d_min64((n2 + 255) / 256, 64));
    k_brute<<<grid, 256, 0, stream>>>(aabb1, n1, aabb2, n2, out_pairs, cap, out_count);
    D3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

}  // extern "C"
