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
//   k_blocks        the sorted boxes are cut into BLOCKS of 32 neighbours (one warp's worth):
//                   64-byte leaf records (exact fp64 box + object index) in sorted order, the
//                   union box of every block, and the block's key (Morton code of its first box)
//   k_hierarchy     Karras 2012 hierarchy over the BLOCKS: one thread per internal node, binary
//                   search on the common prefix of the (unique) keys; k_ropes adds the "rope"
//                   (next node in depth-first order when a subtree is skipped)
//   k_refit         bottom-up AABB union, second arrival proceeds (atomic flags)
//   k_tile<M,SELF>  the query kernel.  A warp owns 32 queries that are neighbours in space; it
//                   walks the block tree ONCE for all of them with the union of their boxes
//                   (stackless: node = overlap ? first child : rope; 32-byte fp32 node records
//                   rounded outward, warp-uniform) and at every block whose box meets the union
//                   it stages the block's 32 exact boxes in shared memory and tests the 32 x 32
//                   tile: each lane compares its own query with the 32 staged boxes (broadcast
//                   reads, the reference's closed-interval fp64 predicate) and collects a 32-bit
//                   hit mask.  Hits are staged per warp and flushed with one reservation on the
//                   output cursor and coalesced 8-byte stores.
//   k_thread<M>     one independent traversal per thread (query sets without spatial order)
//   k_brute         brute force; its pairs are appended with a warp-aggregated atomic
//
// The tree over 1 M boxes has 31 250 leaves (2 MB of node records, L1 / L2 resident), the hierarchy
// and refit kernels work on 32x fewer nodes, and a leaf visit does 1024 box tests in ~500 warp
// instructions instead of one node fetch per box and packet.
#include "d3d_common.cuh"

namespace {

#ifndef BLOCK_LEAVES
#define BLOCK_LEAVES 8  // boxes per block (leaf of the tree); 32 / BLOCK_LEAVES blocks per warp of queries
#endif

struct __align__(16) BvhNode {  // build record of the block tree (exact fp64 boxes)
    double lo[3];
    double hi[3];
    int left, right, rope, parent;
};
static_assert(sizeof(BvhNode) == 64, "node record must be 64 bytes");

struct __align__(16) TNode {  // traversal record of a node of the block tree
    float lo[3];
    float hi[3];
    int left, rope;  // left >= 0: first child; left < 0: leaf = block -left - 1
};
static_assert(sizeof(TNode) == 32, "traversal record must be 32 bytes");
struct __align__(16) LeafRec {  // one box in sorted order: the exact box and its object
    double box[6];              // lo.x hi.x lo.y hi.y lo.z hi.z (the (3,2) layout of the input)
    int obj;
    int pad[3];
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
    unsigned long long *bkeys;    // [nb] key of every block
    int *flags;                   // [nb-1]
    int *range_first;             // [nb-1]
    int *range_last;              // [2nb-1] last block covered by each node
    BvhNode *nodes;               // [2nb-1] block tree, build records
    TNode *tnodes;                // [2nb-1] block tree, traversal records
    LeafRec *leaves;              // [n] boxes in sorted order
    int sort_blocks;
    int nb;                       // number of blocks
};

inline size_t au(size_t x) { return (x + 255) / 256 * 256; }
inline size_t n_blocks_of(int64_t n) { return (size_t)((n > 0 ? n : 1) + BLOCK_LEAVES - 1) / BLOCK_LEAVES; }

inline size_t bvh_ws_bytes(int64_t n) {
    size_t nn = (size_t)(n > 0 ? n : 1);
    size_t nb = n_blocks_of(n);
    size_t sort_blocks = (nn + SORT_TILE - 1) / SORT_TILE;
    return 256 + au(1024 * 6 * 8) + 1024 + 2 * au(nn * 8) + au(256 * sort_blocks * 4) + au(nb * 8) +
           2 * au(nb * 4) + au(2 * nb * 4) + au(2 * nb * sizeof(BvhNode)) + au(2 * nb * sizeof(TNode)) +
           au(nn * sizeof(LeafRec));
}

inline BvhLayout bvh_carve(void *ws, int64_t n) {
    size_t nn = (size_t)(n > 0 ? n : 1);
    size_t nb = n_blocks_of(n);
    BvhLayout L;
    char *p = reinterpret_cast<char *>(ws);
    L.hdr = reinterpret_cast<BvhHeader *>(p); p += 256;
    L.partials = reinterpret_cast<double *>(p); p += au(1024 * 6 * 8);
    L.digit_totals = reinterpret_cast<unsigned *>(p); p += 1024;
    L.keys[0] = reinterpret_cast<unsigned long long *>(p); p += au(nn * 8);
    L.keys[1] = reinterpret_cast<unsigned long long *>(p); p += au(nn * 8);
    L.sort_blocks = (int)((nn + SORT_TILE - 1) / SORT_TILE);
    L.hist = reinterpret_cast<unsigned *>(p); p += au(256 * (size_t)L.sort_blocks * 4);
    L.bkeys = reinterpret_cast<unsigned long long *>(p); p += au(nb * 8);
    L.flags = reinterpret_cast<int *>(p); p += au(nb * 4);
    L.range_first = reinterpret_cast<int *>(p); p += au(nb * 4);
    L.range_last = reinterpret_cast<int *>(p); p += au(2 * nb * 4);
    L.nodes = reinterpret_cast<BvhNode *>(p); p += au(2 * nb * sizeof(BvhNode));
    L.tnodes = reinterpret_cast<TNode *>(p); p += au(2 * nb * sizeof(TNode));
    L.leaves = reinterpret_cast<LeafRec *>(p);
    L.nb = (int)nb;
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

// ---------------------------------------------------------------- blocks + hierarchy
__device__ __forceinline__ int delta(const unsigned long long *keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    return __clzll(keys[i] ^ keys[j]);  // keys are unique (index in the low bits)
}

__device__ __forceinline__ void store_tbox(TNode *t, const double *lo, const double *hi) {
    float4 a = make_float4(__double2float_rd(lo[0]), __double2float_rd(lo[1]), __double2float_rd(lo[2]),
                           __double2float_ru(hi[0]));
    float2 b = make_float2(__double2float_ru(hi[1]), __double2float_ru(hi[2]));
    *reinterpret_cast<float4 *>(t) = a;
    *reinterpret_cast<float2 *>(&t->hi[1]) = b;
}

// Thread i = sorted position i: writes its leaf record; warp b = block b: union box and key of the
// block, the block's leaf record of the tree (build + traversal copy).  256 threads = 8 blocks.
__global__ void __launch_bounds__(256)
k_blocks(const double *__restrict__ aabb, const unsigned long long *__restrict__ keys, int n, int nb,
         LeafRec *leaves, unsigned long long *bkeys, BvhNode *nodes, TNode *tnodes, int *range_last) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    double lo[3] = {1e308, 1e308, 1e308}, hi[3] = {-1e308, -1e308, -1e308};
    unsigned long long key = 0;
    if (i < n) {
        key = keys[i];
        int obj = (int)(unsigned)(key & 0xffffffffull);
        const double2 *b = reinterpret_cast<const double2 *>(aabb + 6 * (int64_t)obj);
        double2 x = __ldg(b), y = __ldg(b + 1), z = __ldg(b + 2);
        double2 *lb = reinterpret_cast<double2 *>(leaves + i);
        lb[0] = x; lb[1] = y; lb[2] = z;
        leaves[i].obj = obj;
        lo[0] = x.x; hi[0] = x.y; lo[1] = y.x; hi[1] = y.y; lo[2] = z.x; hi[2] = z.y;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int off = BLOCK_LEAVES / 2; off > 0; off >>= 1) {  // reduction inside the block's lanes
            lo[k] = fmin(lo[k], __shfl_xor_sync(FULL, lo[k], off));
            hi[k] = fmax(hi[k], __shfl_xor_sync(FULL, hi[k], off));
        }
    int b = i / BLOCK_LEAVES;
    if (lane % BLOCK_LEAVES == 0 && b < nb) {
        bkeys[b] = (key & 0xffffffff00000000ull) | (unsigned long long)(unsigned)b;
        BvhNode *nd = nodes + (nb - 1 + b);
#pragma unroll
        for (int k = 0; k < 3; ++k) { nd->lo[k] = lo[k]; nd->hi[k] = hi[k]; }
        nd->left = -(b + 1);
        nd->right = -1;
        nd->rope = -1;
        if (nb == 1) nd->parent = -1;
        store_tbox(tnodes + (nb - 1 + b), lo, hi);
        tnodes[nb - 1 + b].left = -(b + 1);
        if (nb == 1) tnodes[0].rope = -1;
        range_last[nb - 1 + b] = b;
    }
}

// Karras hierarchy over the nb block keys (thread i < nb - 1 = internal node i).
__global__ void k_hierarchy(const unsigned long long *__restrict__ keys, int n, BvhNode *nodes, TNode *tnodes,
                            int *range_first, int *range_last, int *flags) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
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
                        const int *__restrict__ range_last) {
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
    tnodes[v].rope = rope;
}

// Bottom-up refit: every leaf thread climbs, the second thread to arrive at a node merges the
// two child boxes and goes on.  A node whose leaf range lies inside the 256 leaves of this
// block can only be reached by threads of this block: its arrival flag lives in shared memory
// and block-scope fences order the box stores; only the nodes that span blocks pay for
// device-scope fences and global atomics.
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

struct Tree {
    const TNode *__restrict__ tnodes;
    const LeafRec *__restrict__ leaves;
    int nb;  // blocks; leaf node of block b = nb - 1 + b
    int n;   // boxes
};

// aabb_tree.py:520-527 (closed intervals) on exact boxes in the (3,2) layout
__device__ __forceinline__ bool boxes_overlap(double2 ax, double2 ay, double2 az, double2 bx, double2 by,
                                              double2 bz) {
    return ax.x <= bx.y && ax.y >= bx.x && ay.x <= by.y && ay.y >= by.x && az.x <= bz.y && az.y >= bz.x;
}

enum { Q_APPEND = 0, Q_COUNT = 1, Q_FILL = 2 };

// Fused all-gather: when `bufs` is set, a flushed chunk is stored into the pair buffer of EVERY
// GPU of the job (peer pointers over NVLink / NVSwitch, this GPU's own buffer included) at this
// rank's segment, so the list is complete everywhere when the traversals end and the transfer
// overlaps the walk chunk by chunk.
struct PeerOut {
    int32_t *const *bufs;  // device array of n pointers (symmetric allocation), or nullptr
    int n, self;
    int64_t seg_base;      // first pair slot of this rank's segment
};

__device__ __forceinline__ void flush_stage(const int2 *stage, int staged, int lane, int32_t *out_pairs,
                                            int64_t cap, unsigned long long *cursor, const PeerOut &peers) {
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(cursor, (unsigned long long)staged);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (peers.bufs == nullptr) {
        for (int i = lane; i < staged; i += 32)
            if ((int64_t)(base + i) < cap) reinterpret_cast<int2 *>(out_pairs)[base + i] = stage[i];
    } else {
        for (int q = 0; q < peers.n; ++q) {  // start behind myself so that the ranks do not all hit one peer
            int p = peers.self + 1 + q;
            if (p >= peers.n) p -= peers.n;
            int2 *dst = reinterpret_cast<int2 *>(peers.bufs[p]) + peers.seg_base;
            for (int i = lane; i < staged; i += 32)
                if ((int64_t)(base + i) < cap) dst[base + i] = stage[i];
        }
    }
}
#ifndef TILE_STAGE
#define TILE_STAGE (512 + BLOCK_LEAVES * 32)  // staged pairs per warp: flushed when less than one full tile is free
#endif
#define TILE_WARPS 4

// The query kernel (see the file header).  M: Q_APPEND = one pass, hits appended through the
// per-warp stage (pair order depends on scheduling, *cursor = exact count even when cap is
// too small); Q_COUNT / Q_FILL = the two passes of the ordered query (per-query counts, then
// every lane writes its pairs behind its own offset: queries in processing order, boxes in
// sorted order).  SELF: the queries are the tree's own boxes and every unordered pair is wanted
// once, without (i, i): warp w of part p owns block (CTA * n_parts + p) * 4 + w, starts AT its own
// block (the depth-first order of the tree is the sorted order, so everything behind a block is
// reached by following ropes from it) and keeps pairs whose partner lies behind the query.
template <int M, bool SELF>
__global__ void __launch_bounds__(TILE_WARPS * 32)
k_tile(Tree T, const BvhHeader *hdr, const double *__restrict__ query, const int32_t *__restrict__ order,
       int part, int n_parts, int64_t n_query, unsigned *__restrict__ counts,
       const unsigned long long *__restrict__ offsets, int32_t *out_pairs, int64_t cap,
       unsigned long long *cursor, unsigned long long *visits, PeerOut peers) {
    extern __shared__ double tile_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned FULL = 0xffffffffu;
    // per warp: 32 staged boxes (6 doubles each) + 32 object indices + the pair stage
    double *sbox = tile_smem + wid * (BLOCK_LEAVES * 6);
    int *sobj = reinterpret_cast<int *>(tile_smem + TILE_WARPS * BLOCK_LEAVES * 6) + wid * BLOCK_LEAVES;
    int2 *stage = reinterpret_cast<int2 *>(tile_smem + TILE_WARPS * BLOCK_LEAVES * 6 + TILE_WARPS * BLOCK_LEAVES / 2) +
                  (M == Q_APPEND ? wid * TILE_STAGE : 0);
    const int64_t warp_id = SELF ? ((int64_t)blockIdx.x * n_parts + part) * TILE_WARPS + wid
                                 : (int64_t)blockIdx.x * TILE_WARPS + wid;
    const int64_t t = warp_id * 32 + lane;
    const bool valid = t < n_query;
    int qi = 0;
    double2 qx = make_double2(1e308, -1e308), qy = qx, qz = qx;  // empty box: never overlaps
    if (valid) {
        const double2 *qb;
        if (SELF) {
            qi = __ldg(&T.leaves[t].obj);
            qb = reinterpret_cast<const double2 *>(T.leaves + t);
        } else {
            qi = order ? order[t] : (int)t;
            qb = reinterpret_cast<const double2 *>(query + 6 * (int64_t)qi);
        }
        qx = __ldg(qb); qy = __ldg(qb + 1); qz = __ldg(qb + 2);
    }
    // this lane's box rounded outward to fp32; the warp descends where ANY of its queries overlaps
    // (a union box would blow up for the warps that straddle a jump of the Morton curve)
    float mlo[3] = {__double2float_rd(qx.x), __double2float_rd(qy.x), __double2float_rd(qz.x)};
    float mhi[3] = {__double2float_ru(qx.y), __double2float_ru(qy.y), __double2float_ru(qz.y)};
    if (!valid) { mlo[0] = mlo[1] = mlo[2] = 3.0e38f; mhi[0] = mhi[1] = mhi[2] = -3.0e38f; }
    const bool warp_valid = __any_sync(FULL, valid);
    const int my_block = (int)(t / BLOCK_LEAVES);  // SELF: the block this lane's query sits in
    int node = -1;
    if (hdr->n > 0 && warp_valid) node = SELF ? T.nb - 1 + (int)(warp_id * (32 / BLOCK_LEAVES)) : hdr->root;
    int staged = 0;  // warp-uniform
    unsigned n_hits = 0, n_nodes = 0, n_tiles = 0;
    unsigned long long pos = (M == Q_FILL && valid) ? offsets[t] : 0ull;
    while (node >= 0) {
        const float4 *p = reinterpret_cast<const float4 *>(T.tnodes + node);
        float4 a = __ldg(p), b = __ldg(p + 1);  // lo.x lo.y lo.z hi.x | hi.y hi.z left rope (same address: broadcast)
        const bool mine = a.x <= mhi[0] && a.w >= mlo[0] && a.y <= mhi[1] && b.x >= mlo[1] && a.z <= mhi[2] && b.y >= mlo[2];
        const bool ov = __any_sync(FULL, mine);
        const int left = __float_as_int(b.z), rope = __float_as_int(b.w);
        ++n_nodes;
        if (ov && left < 0) {
            // ---- 32 x 32 tile: this warp's queries against the boxes of block blk
            const int blk = -left - 1;
            const int first = blk * BLOCK_LEAVES;
            const int cnt = min(BLOCK_LEAVES, T.n - first);
            __syncwarp();
            if (lane < cnt) {
                const double2 *lb = reinterpret_cast<const double2 *>(T.leaves + first + lane);
                double2 x = __ldg(lb), y = __ldg(lb + 1), z = __ldg(lb + 2);
                double2 *sb = reinterpret_cast<double2 *>(sbox + 6 * lane);
                sb[0] = x; sb[1] = y; sb[2] = z;
                sobj[lane] = __ldg(&T.leaves[first + lane].obj);
            }
            __syncwarp();
            unsigned mask = 0;
            if (mine) {
#pragma unroll 4
                for (int k = 0; k < cnt; ++k) {
                    const double2 *sb = reinterpret_cast<const double2 *>(sbox + 6 * k);
                    if (boxes_overlap(sb[0], sb[1], sb[2], qx, qy, qz)) mask |= 1u << k;
                }
                if (SELF) {  // partners behind the query only
                    if (blk < my_block) mask = 0;
                    else if (blk == my_block) mask &= ~((2u << (lane % BLOCK_LEAVES)) - 1u);
                }
            }
            ++n_tiles;
            const int c = __popc(mask);
            if (M == Q_COUNT) {
                n_hits += c;
            } else if (M == Q_FILL) {
                for (unsigned m = mask; m; m &= m - 1) {
                    int k = __ffs(m) - 1;
                    if ((int64_t)pos < cap) reinterpret_cast<int2 *>(out_pairs)[pos] = make_int2(sobj[k], qi);
                    ++pos;
                }
            } else {
                int incl = c;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    int y = __shfl_up_sync(FULL, incl, off);
                    if (lane >= off) incl += y;
                }
                const int total = __shfl_sync(FULL, incl, 31);
                int at = staged + incl - c;
                for (unsigned m = mask; m; m &= m - 1) {
                    int k = __ffs(m) - 1, o = sobj[k];
                    stage[at++] = SELF ? make_int2(min(o, qi), max(o, qi)) : make_int2(o, qi);
                }
                staged += total;
                if (staged > TILE_STAGE - BLOCK_LEAVES * 32) {  // room for one more full tile
                    __syncwarp();
                    flush_stage(stage, staged, lane, out_pairs, cap, cursor, peers);
                    staged = 0;
                    __syncwarp();  // the stage is rewritten from slot 0
                }
            }
        }
        node = (ov && left >= 0) ? left : rope;
    }
    if (M == Q_APPEND && staged) {
        __syncwarp();
        flush_stage(stage, staged, lane, out_pairs, cap, cursor, peers);
    }
    if (M == Q_COUNT && valid) counts[t] = n_hits;
    // measurement: 32-byte node records + 64-byte leaf records fetched, in units of 32 bytes
    if (visits && M != Q_FILL && lane == 0 && warp_valid)
        atomicAdd(visits, (unsigned long long)n_nodes + (unsigned long long)n_tiles * (2 * BLOCK_LEAVES));
}

// One independent traversal per thread, for query sets without spatial order (a warp's union box
// would cover everything).  At a block the thread tests the block's boxes itself.
template <int M>
__global__ void __launch_bounds__(128)
k_thread(Tree T, const BvhHeader *hdr, const double *__restrict__ query, const int32_t *__restrict__ order,
         int64_t n_query, unsigned *__restrict__ counts, const unsigned long long *__restrict__ offsets,
         int32_t *out_pairs, int64_t cap, unsigned long long *cursor, unsigned long long *visits) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n_query) return;
    int qi = order ? order[t] : (int)t;
    const double2 *qb = reinterpret_cast<const double2 *>(query + 6 * (int64_t)qi);
    double2 qx = __ldg(qb), qy = __ldg(qb + 1), qz = __ldg(qb + 2);
    const float lox = __double2float_rd(qx.x), hix = __double2float_ru(qx.y), loy = __double2float_rd(qy.x),
                hiy = __double2float_ru(qy.y), loz = __double2float_rd(qz.x), hiz = __double2float_ru(qz.y);
    int node = hdr->n > 0 ? hdr->root : -1;
    unsigned n_hits = 0, n_visited = 0;
    unsigned long long pos = M == Q_FILL ? offsets[t] : 0ull;
    while (node >= 0) {
        const float4 *p = reinterpret_cast<const float4 *>(T.tnodes + node);
        float4 a = __ldg(p), b = __ldg(p + 1);
        const bool ov = a.x <= hix && a.w >= lox && a.y <= hiy && b.x >= loy && a.z <= hiz && b.y >= loz;
        const int left = __float_as_int(b.z), rope = __float_as_int(b.w);
        ++n_visited;
        if (ov && left < 0) {
            const int first = (-left - 1) * BLOCK_LEAVES;
            const int cnt = min(BLOCK_LEAVES, T.n - first);
            n_visited += 2 * cnt;
#pragma unroll 1
            for (int k = 0; k < cnt; ++k) {
                const double2 *lb = reinterpret_cast<const double2 *>(T.leaves + first + k);
                if (boxes_overlap(__ldg(lb), __ldg(lb + 1), __ldg(lb + 2), qx, qy, qz)) {
                    int o = __ldg(&T.leaves[first + k].obj);
                    if (M == Q_COUNT) ++n_hits;
                    else if (M == Q_FILL) {
                        if ((int64_t)pos < cap) reinterpret_cast<int2 *>(out_pairs)[pos] = make_int2(o, qi);
                        ++pos;
                    } else append_pair(o, qi, out_pairs, cap, cursor);
                }
            }
        }
        node = (ov && left >= 0) ? left : rope;
    }
    if (M == Q_COUNT) counts[t] = n_hits;
    if (visits && M != Q_FILL) atomicAdd(visits, (unsigned long long)n_visited);
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

// after a fused traversal: every GPU learns this rank's pair count
__global__ void k_publish_count(const unsigned long long *cursor, unsigned long long *const *peer_counts, int n,
                                int self) {
    int p = threadIdx.x;
    if (p < n) peer_counts[p][self] = *cursor;
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

static inline Tree tree_of(const BvhLayout &L, int64_t n) {
    Tree T;
    T.tnodes = L.tnodes; T.leaves = L.leaves; T.nb = L.nb; T.n = (int)n;
    return T;
}

static const size_t TILE_SMEM_FIXED = (size_t)TILE_WARPS * BLOCK_LEAVES * 6 * 8 + (size_t)TILE_WARPS * BLOCK_LEAVES * 4;
static const size_t TILE_SMEM_APPEND = TILE_SMEM_FIXED + (size_t)TILE_WARPS * TILE_STAGE * 8;

template <int M, bool SELF>
static int launch_tile(const BvhLayout &L, int64_t n, const double *query, const int32_t *order, int part,
                       int n_parts, int64_t n_query, unsigned *counts, const unsigned long long *offsets,
                       int32_t *out_pairs, int64_t cap, unsigned long long *cursor, unsigned long long *visits,
                       cudaStream_t stream, PeerOut peers = PeerOut{nullptr, 0, 0, 0}) {
    const size_t smem = M == Q_APPEND ? TILE_SMEM_APPEND : TILE_SMEM_FIXED;
    static bool attr_set[64] = {false};  // per instance of this template, per device
    int dev = 0;
    D3D_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        D3D_CUDA_CHECK(cudaFuncSetAttribute(k_tile<M, SELF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    int64_t all_ctas = (n_query + TILE_WARPS * 32 - 1) / (TILE_WARPS * 32);
    int64_t ctas = SELF ? (all_ctas - part + n_parts - 1) / n_parts : all_ctas;
    if (ctas <= 0) return 0;
    k_tile<M, SELF><<<(unsigned)ctas, TILE_WARPS * 32, smem, stream>>>(tree_of(L, n), L.hdr, query, order, part, n_parts,
                                                                      n_query, counts, offsets, out_pairs, cap,
                                                                      cursor, visits, peers);
    D3D_CUDA_CHECK(cudaGetLastError());
    return 0;
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
    unsigned nblk = (unsigned)((n + 255) / 256);
    k_morton<<<nblk, 256, 0, stream>>>(aabb, n, L.hdr, L.keys[0]);
    int cur = 0;
    for (int shift = 32; shift < 64; shift += 8) {
        k_sort_hist<<<L.sort_blocks, SORT_THREADS, 0, stream>>>(L.keys[cur], n, shift, L.hist, L.sort_blocks);
        k_sort_scan<<<32, 256, 0, stream>>>(L.hist, L.sort_blocks, L.digit_totals);
        k_sort_scatter<<<L.sort_blocks, SORT_THREADS, 0, stream>>>(L.keys[cur], L.keys[cur ^ 1], n, shift,
                                                                  L.hist, L.sort_blocks, L.digit_totals);
        cur ^= 1;
    }
    // 4 passes: sorted keys are back in keys[0]
    const int nb = L.nb;
    k_blocks<<<nblk, 256, 0, stream>>>(aabb, L.keys[0], (int)n, nb, L.leaves, L.bkeys, L.nodes, L.tnodes, L.range_last);
    if (nb > 1) {
        unsigned hb = (unsigned)((nb + 255) / 256);
        k_hierarchy<<<hb, 256, 0, stream>>>(L.bkeys, nb, L.nodes, L.tnodes, L.range_first, L.range_last, L.flags);
        k_ropes<<<(unsigned)((2 * nb - 1 + 255) / 256), 256, 0, stream>>>(L.bkeys, nb, L.nodes, L.tnodes, L.range_last);
        k_refit<<<hb, 256, 0, stream>>>(nb, L.nodes, L.tnodes, L.flags, L.range_first, L.range_last);
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
    if (packet) {
        int rc = launch_tile<Q_COUNT, false>(L, n, query, order, 0, 1, n_query, Q.counts, nullptr, nullptr, 0,
                                             nullptr, out_visits, stream);
        if (rc) return rc;
    } else {
        k_thread<Q_COUNT><<<(unsigned)((n_query + 127) / 128), 128, 0, stream>>>(
            tree_of(L, n), L.hdr, query, order, n_query, Q.counts, nullptr, nullptr, 0, nullptr, out_visits);
    }
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
        return launch_tile<Q_FILL, false>(L, n, query, order, 0, 1, n_query, nullptr, Q.offsets, out_pairs, cap,
                                          nullptr, nullptr, stream);
    k_thread<Q_FILL><<<(unsigned)((n_query + 127) / 128), 128, 0, stream>>>(
        tree_of(L, n), L.hdr, query, order, n_query, nullptr, Q.offsets, out_pairs, cap, nullptr, nullptr);
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
    if (packet < 0 || packet > 32) return d3d_set_error("d3d_bvh_overlap: packet must be 0 (per thread) or 1 (warp tiles)");
    BvhLayout L = bvh_carve(const_cast<void *>(workspace), n);
    if (packet)
        return launch_tile<Q_APPEND, false>(L, n, query, order, 0, 1, n_query, nullptr, nullptr, out_pairs, cap,
                                            out_count, out_visits, stream);
    k_thread<Q_APPEND><<<(unsigned)((n_query + 127) / 128), 128, 0, stream>>>(
        tree_of(L, n), L.hdr, query, order, n_query, nullptr, nullptr, out_pairs, cap, out_count, out_visits);
    D3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

/* the tree against its own boxes: every unordered pair once, as (smaller, larger) object index,
 * no (i, i).  Part `part` of `n_parts` (one per GPU; 0 of 1 = everything) takes the 128-box
 * groups part, part + n_parts, ... of the sorted order. */
int d3d_bvh_overlap_self(const void *workspace, int64_t n, int part, int n_parts, int packet,
                         int32_t *out_pairs, int64_t cap, unsigned long long *out_count,
                         unsigned long long *out_visits, void *stream_) {
    (void)packet;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!workspace || !out_count) return d3d_set_error("d3d_bvh_overlap_self: null argument");
    D3D_CUDA_CHECK(cudaMemsetAsync(out_count, 0, sizeof(unsigned long long), stream));
    if (out_visits) D3D_CUDA_CHECK(cudaMemsetAsync(out_visits, 0, sizeof(unsigned long long), stream));
    if (n == 0) return 0;
    if (n_parts < 1 || part < 0 || part >= n_parts) return d3d_set_error("d3d_bvh_overlap_self: part out of range");
    if (cap > 0 && !out_pairs) return d3d_set_error("d3d_bvh_overlap_self: null output");
    BvhLayout L = bvh_carve(const_cast<void *>(workspace), n);
    return launch_tile<Q_APPEND, true>(L, n, nullptr, nullptr, part, n_parts, n, nullptr, nullptr, out_pairs, cap,
                                       out_count, out_visits, stream);
}

/* d3d_bvh_overlap_self fused with the all-gather of the ranks' pair lists: see include/d3d_b200.h */
int d3d_bvh_overlap_self_gather(const void *workspace, int64_t n, int part, int n_parts,
                                int32_t *const *peer_pairs, int64_t segment_cap,
                                unsigned long long *const *peer_counts, unsigned long long *local_count,
                                unsigned long long *out_visits, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!workspace || !peer_pairs || !peer_counts || !local_count)
        return d3d_set_error("d3d_bvh_overlap_self_gather: null argument");
    if (n_parts < 1 || n_parts > 32 || part < 0 || part >= n_parts)
        return d3d_set_error("d3d_bvh_overlap_self_gather: part out of range (at most 32 GPUs)");
    D3D_CUDA_CHECK(cudaMemsetAsync(local_count, 0, sizeof(unsigned long long), stream));
    if (out_visits) D3D_CUDA_CHECK(cudaMemsetAsync(out_visits, 0, sizeof(unsigned long long), stream));
    if (n > 0) {
        BvhLayout L = bvh_carve(const_cast<void *>(workspace), n);
        PeerOut peers;
        peers.bufs = peer_pairs; peers.n = n_parts; peers.self = part; peers.seg_base = (int64_t)part * segment_cap;
        int rc = launch_tile<Q_APPEND, true>(L, n, nullptr, nullptr, part, n_parts, n, nullptr, nullptr, nullptr,
                                             segment_cap, local_count, out_visits, stream, peers);
        if (rc) return rc;
    }
    k_publish_count<<<1, 32, 0, stream>>>(local_count, peer_counts, n_parts, part);
    D3D_CUDA_CHECK(cudaGetLastError());
    return 0;
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
