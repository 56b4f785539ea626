// bvh_build.cu — LBVH construction on the device: Morton-3D, stable LSD radix sort, Karras
// internal nodes with the reference's duplicate-key rule, atomic bottom-up AABB refit and a
// parent-walk pre-order flatten.  Replaces Bvh.setup_data_gpu (accel/LBvh.py:192-226):
//   build_morton_3d (:318-336), radix_sort_host/blelloch_* (:55-72,339-386; 450-1170 launches in
//   the reference), build_lbvh (:389-450), the gen_aabb host loop (:206-218, one D->H sync per
//   sweep) and the host-side Python recursion flatten_tree (:138-173).
// Here: 1 + 12 + 1 + 1 + 1 launches, no host round trip until the final status read.
#include "ctx.h"
#include "common.cuh"

#define SORT_THREADS 256
#define SORT_ROUNDS 8
#define SORT_CHUNK (SORT_THREADS * SORT_ROUNDS)
#define RADIX 256

// ------------------------------------------------------------------ Morton (accel/LBvh.py:318-336)
__global__ void k_morton(const float* __restrict__ vertex, const int* __restrict__ prim, const float* __restrict__ shape,
                         int n, V3 bmin, V3 bmax, int* __restrict__ keys, int* __restrict__ vals, int* __restrict__ unsorted) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int type = prim[i * 3], vid = prim[i * 3 + 1];
    int code;
    if (type == TR_PRIM_TRI) {
        const float* p = vertex + (size_t)vid * 9;
        V3 v0 = mk3(p[0], p[1], p[2]), v1 = mk3(p[9], p[10], p[11]), v2 = mk3(p[18], p[19], p[20]);
        V3 c = ((v1 + v2) + v0) * (1.0f / 3.0f);
        V3 d = bmax - bmin, q = c - bmin;
        code = morton3d(q.x / d.x, q.y / d.y, q.z / d.z);
    } else {
        const float* s = shape + (size_t)vid * 10;          // (type, pos.x, pos.y) un-normalised, LBvh.py:333-335
        code = morton3d(s[0], s[1], s[2]);
    }
    keys[i] = code; vals[i] = i;
    unsorted[i * 2] = code; unsorted[i * 2 + 1] = i;
}

// ------------------------------------------------------------------ radix sort, 8 bits per pass
// Stable LSD sort of (code, prim) by the low 30 bits -> same permutation as the reference's 30
// one-bit Blelloch passes (ties keep the original primitive order).
__global__ void k_sort_hist(const int* __restrict__ keys, int n, int shift, int nblocks, int* __restrict__ hist) {
    __shared__ int h[RADIX];
    h[threadIdx.x] = 0;
    __syncthreads();
    int base = blockIdx.x * SORT_CHUNK;
#pragma unroll
    for (int r = 0; r < SORT_ROUNDS; ++r) {
        int i = base + r * SORT_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[((unsigned)keys[i] >> shift) & (RADIX - 1)], 1);
    }
    __syncthreads();
    hist[threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];   // digit-major
}

// single-block exclusive scan over RADIX*nblocks counters
__global__ void k_sort_scan(int* __restrict__ hist, int total) {
    __shared__ int warp_sums[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int T = blockDim.x;
    for (int base = 0; base < total; base += T) {
        int i = base + threadIdx.x;
        int v = (i < total) ? hist[i] : 0;
        int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) warp_sums[w] = x;
        __syncthreads();
        if (w == 0) {
            int s = (lane < (T >> 5)) ? warp_sums[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
            warp_sums[lane] = s;
        }
        __syncthreads();
        int excl = x - v + (w > 0 ? warp_sums[w - 1] : 0) + carry;
        if (i < total) hist[i] = excl;
        __syncthreads();
        if (threadIdx.x == T - 1) carry = excl + v;
        __syncthreads();
    }
}

__global__ void k_sort_scatter(const int* __restrict__ keys, const int* __restrict__ vals, int n, int shift, int nblocks,
                               const int* __restrict__ hist, int* __restrict__ okeys, int* __restrict__ ovals) {
    __shared__ int base[RADIX];
    __shared__ int whist[SORT_THREADS / 32][RADIX];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    base[threadIdx.x] = hist[threadIdx.x * nblocks + blockIdx.x];
#pragma unroll
    for (int k = 0; k < SORT_THREADS / 32; ++k) whist[k][threadIdx.x] = 0;
    __syncthreads();
    int cbase = blockIdx.x * SORT_CHUNK;
    for (int r = 0; r < SORT_ROUNDS; ++r) {
        int i = cbase + r * SORT_THREADS + threadIdx.x;
        bool valid = i < n;
        int key = valid ? keys[i] : 0, val = valid ? vals[i] : 0;
        unsigned digit = valid ? (((unsigned)key >> shift) & (RADIX - 1)) : 0xFFFFFFFFu;
        unsigned peers = __match_any_sync(0xffffffffu, digit);
        int rank = __popc(peers & ((1u << lane) - 1u));
        if (valid && rank == 0) whist[w][digit] = __popc(peers);
        __syncthreads();
        if (valid) {
            int off = base[digit] + rank;
            for (int k = 0; k < w; ++k) off += whist[k][digit];
            okeys[off] = key; ovals[off] = val;
        }
        __syncthreads();
        int add = 0;
#pragma unroll
        for (int k = 0; k < SORT_THREADS / 32; ++k) { add += whist[k][threadIdx.x]; whist[k][threadIdx.x] = 0; }
        base[threadIdx.x] += add;
        __syncthreads();
    }
}

// ------------------------------------------------------------------ Karras (accel/LBvh.py:229-314,428-450)
__device__ __forceinline__ void determine_range(const int* __restrict__ code, int n, int idx, int& r0, int& r1) {
    r0 = 0; r1 = n - 1;
    if (idx == 0) return;
    int self = code[idx], lc = code[idx - 1], rc = code[idx + 1];
    if (lc == self && rc == self) {
        // duplicate rule of the reference (:240-251): range = [idx, end of the run of equal codes]
        r0 = idx;
        while (idx < n - 1) {
            idx += 1;
            if (idx >= n - 1) break;
            if (code[idx] != code[idx + 1]) break;
        }
        r1 = idx;
        return;
    }
    int Ld = common_upper_bits(self, lc), Rd = common_upper_bits(self, rc);
    int d = (Rd > Ld) ? 1 : -1;
    int dmin = min(Ld, Rd);
    int lmax = 2, delta = -1, it = idx + d * lmax;
    if (it >= 0 && it < n) delta = common_upper_bits(self, code[it]);
    while (delta > dmin) {
        lmax <<= 1; it = idx + d * lmax; delta = -1;
        if (it >= 0 && it < n) delta = common_upper_bits(self, code[it]);
    }
    int l = 0;
    for (int t = lmax >> 1; t > 0; t >>= 1) {
        it = idx + (l + t) * d; delta = -1;
        if (it >= 0 && it < n) delta = common_upper_bits(self, code[it]);
        if (delta > dmin) l += t;
    }
    r0 = idx; r1 = idx + l * d;
    if (d < 0) { int t = r0; r0 = r1; r1 = t; }
}

__device__ __forceinline__ int find_split(const int* __restrict__ code, int first, int last) {
    int fc = code[first], lc = code[last], split = first;
    if (fc != lc) {
        int dn = common_upper_bits(fc, lc), stride = last - first;
        while (true) {
            stride = (stride + 1) >> 1;
            int middle = split + stride;
            if (middle < last && common_upper_bits(fc, code[middle]) > dn) split = middle;
            if (stride <= 1) break;
        }
    }
    return split;
}

__global__ void k_karras(const int* __restrict__ code, int n, int* __restrict__ left, int* __restrict__ right, int* __restrict__ parent,
                         int* __restrict__ first) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int r0, r1; determine_range(code, n, i, r0, r1);
    int split = find_split(code, r0, r1);
    int l = split, r = split + 1;
    if (min(r0, r1) == split) l += n - 1;
    if (max(r0, r1) == split + 1) r += n - 1;
    left[i] = l; right[i] = r;
    parent[l] = i; parent[r] = i;
    first[i] = min(r0, r1);                            // sorted position of the node's first leaf
}

// ------------------------------------------------------------------ leaf boxes + atomic bottom-up refit
// Replaces the level-synchronous gen_aabb relaxation (accel/LBvh.py:453-468 + host loop :206-218).
// The fixed point is identical: boxes are exact min/max unions.
__global__ void k_refit(const float* __restrict__ vertex, const int* __restrict__ prim, const float* __restrict__ shape,
                        const int* __restrict__ sorted_prim, int n, const int* __restrict__ left, const int* __restrict__ right,
                        const int* __restrict__ parent, float* boxes, int* leafcount, int* flag, int* status, int* sneed) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int node = n - 1 + k;
    int pi = sorted_prim[k];
    int type = prim[pi * 3], vi = prim[pi * 3 + 1];
    V3 mn = mk3(0, 0, 0), mx = mk3(0, 0, 0);
    if (type == TR_PRIM_TRI) {
        const float* p = vertex + (size_t)vi * 9;
        V3 a = mk3(p[0], p[1], p[2]), b = mk3(p[9], p[10], p[11]), c = mk3(p[18], p[19], p[20]);
        mn = mk3(fminf(fminf(a.x, b.x), c.x), fminf(fminf(a.y, b.y), c.y), fminf(fminf(a.z, b.z), c.z));
        mx = mk3(fmaxf(fmaxf(a.x, b.x), c.x), fmaxf(fmaxf(a.y, b.y), c.y), fmaxf(fmaxf(a.z, b.z), c.z));
    } else {
        const float* s = shape + (size_t)vi * 10;
        if ((int)s[0] == TR_SHAPE_SPHERE) {
            float r = s[4]; V3 c = mk3(s[1], s[2], s[3]);
            mn = c + mk3(-r, -r, -r); mx = c + mk3(r, r, r);
        }
    }
    float* bx = boxes + (size_t)node * 6;
    bx[0] = mn.x; bx[1] = mn.y; bx[2] = mn.z; bx[3] = mx.x; bx[4] = mx.y; bx[5] = mx.z;
    leafcount[node] = 1;
    sneed[node] = 0;
    __threadfence();
    int cur = parent[node];
    while (cur >= 0) {
        int old = atomicAdd(&flag[cur], 1);
        if (old == 0) break;                       // first child to arrive: sibling will finish
        __threadfence();
        int l = left[cur], r = right[cur];
        const float* a = boxes + (size_t)l * 6; const float* b = boxes + (size_t)r * 6;
        float o[6];
#pragma unroll
        for (int q = 0; q < 3; ++q) { o[q] = fminf(__ldcg(a + q), __ldcg(b + q)); o[3 + q] = fmaxf(__ldcg(a + 3 + q), __ldcg(b + 3 + q)); }
        float* c = boxes + (size_t)cur * 6;
#pragma unroll
        for (int q = 0; q < 6; ++q) c[q] = o[q];
        leafcount[cur] = __ldcg(leafcount + l) + __ldcg(leafcount + r);
        // traversal-stack entries the sub-tree needs (trace.cuh: a leaf child is taken before an internal one, the other child
        // is pushed): two internal children 1 + max, one leaf child max(1, S(internal child)), two leaves 1
        const int sl = __ldcg(sneed + l), sr = __ldcg(sneed + r);
        const int need = (l < n - 1 && r < n - 1) ? 1 + max(sl, sr) : max(1, max(sl, sr));
        sneed[cur] = need;
        if (cur == 0) status[1] = need;
        atomicAdd(status, 1);
        __threadfence();
        cur = parent[cur];
    }
}

// ------------------------------------------------------------------ pre-order flatten by parent walk
// pre(x) = #ancestors + sum over ancestors entered through their right child of (2*leaves(left)-1);
// same left-first order as the host recursion flatten_tree (accel/LBvh.py:138-161).
__global__ void k_flatten(const float* __restrict__ vertex, const int* __restrict__ prim, const float* __restrict__ shape,
                          const int* __restrict__ sorted_prim, int n, const int* __restrict__ left, const int* __restrict__ right,
                          const int* __restrict__ parent, const float* __restrict__ boxes, const int* __restrict__ leafcount,
                          int* __restrict__ pre_out, TrNode* __restrict__ nodes, TrLeaf* __restrict__ leaves, int* __restrict__ leaf_of_prim) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int nn = 2 * n - 1;
    if (x >= nn) return;
    int pre = 0, c = x, p;
    while ((p = parent[c]) >= 0) {
        pre += 1;
        if (right[p] == c) pre += 2 * leafcount[left[p]] - 1;
        c = p;
    }
    pre_out[x] = pre;
    const bool leaf = x >= n - 1;
    {   // the reference's pre-order node (exact boxes), for process_normal's point queries
        const float* b = boxes + (size_t)x * 6;
        const int lc = leafcount[x];
        const int link = leaf ? -((x - (n - 1)) + 1) : (pre + 2 * leafcount[left[x]]);
        TrNode nd;
        nd.lo = make_float4(b[0], b[1], b[2], __int_as_float(pre + 2 * lc - 1));
        nd.hi = make_float4(b[3], b[4], b[5], __int_as_float(link));
        nodes[pre] = nd;
    }
    if (leaf) {
        int k = x - (n - 1), pi = sorted_prim[k];
        int type = prim[pi * 3], vi = prim[pi * 3 + 1], mat = prim[pi * 3 + 2];
        leaf_of_prim[pi] = k;
        TrLeaf lf;
        if (type == TR_PRIM_TRI) {
            const float* q = vertex + (size_t)vi * 9;
            V3 v0 = mk3(q[0], q[1], q[2]), v1 = mk3(q[9], q[10], q[11]), v2 = mk3(q[18], q[19], q[20]);
            V3 e1 = v1 - v0, e2 = v2 - v0;                   // Scene.py:614-615
            lf.a = make_float4(v0.x, v0.y, v0.z, __int_as_float(pi));
            lf.b = make_float4(e1.x, e1.y, e1.z, __int_as_float(0));
            lf.c = make_float4(e2.x, e2.y, e2.z, __int_as_float(mat));
        } else {
            const float* s = shape + (size_t)vi * 10;
            int kind = ((int)s[0] == TR_SHAPE_SPHERE) ? 1 : 2;
            lf.a = make_float4(s[1], s[2], s[3], __int_as_float(pi));
            lf.b = make_float4(s[4], 0.0f, 0.0f, __int_as_float(kind));
            lf.c = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(mat));
        }
        leaves[k] = lf;
    }
}

// ------------------------------------------------------------------ render-kernel nodes (TrNode2)
// Breadth-first top of the tree: the first min(TR_TOP_MAX, n-1) internal nodes in level order (one block, level by level with
// a block scan, so the order is deterministic).  top[i] = build-order id of the i-th node.
__global__ void __launch_bounds__(1024) k_top_bfs(int n, const int* __restrict__ left, const int* __restrict__ right, int* __restrict__ top,
                                                   int* __restrict__ status) {
    __shared__ int q[TR_TOP_MAX];
    __shared__ int scan[1024];
    __shared__ int s_cnt;
    const int t = threadIdx.x, nint = n - 1;
    if (t == 0) { q[0] = 0; s_cnt = 1; }
    __syncthreads();
    int lvl_begin = 0, lvl_end = 1;
    while (lvl_begin < lvl_end && lvl_end < TR_TOP_MAX && lvl_end < nint) {
        int c0 = -1, c1 = -1;
        if (lvl_begin + t < lvl_end) {
            const int x = q[lvl_begin + t];
            const int l = left[x], r = right[x];
            if (l < nint) c0 = l;
            if (r < nint) c1 = r;
        }
        const int mine = (c0 >= 0) + (c1 >= 0);
        scan[t] = mine;
        __syncthreads();
        for (int off = 1; off < 1024; off <<= 1) {             // inclusive Hillis-Steele scan
            int v = (t >= off) ? scan[t - off] : 0;
            __syncthreads();
            scan[t] += v;
            __syncthreads();
        }
        int pos = lvl_end + scan[t] - mine;
        if (c0 >= 0) { if (pos < TR_TOP_MAX) q[pos] = c0; ++pos; }
        if (c1 >= 0) { if (pos < TR_TOP_MAX) q[pos] = c1; }
        const int total = scan[1023];
        __syncthreads();
        lvl_begin = lvl_end; lvl_end = min(lvl_end + total, TR_TOP_MAX);
    }
    const int cnt = min(lvl_end, nint);
    if (t < cnt) top[t] = q[t];
    if (t == 0) status[2] = cnt;
}

// position of internal node x (build id) in the TrNode2 array: breadth-first position for the top nodes, otherwise
// cnt + (pre-order rank among internal nodes) - (top nodes with a smaller rank).
__device__ __forceinline__ int node2_index(int x, const int* __restrict__ irank, const int* __restrict__ top_sorted_rank, const int* __restrict__ top_pos, int cnt) {
    const int rk = irank[x];
    int lo = 0, hi = cnt;                      // lower_bound over the top nodes' ranks
    while (lo < hi) { int mid = (lo + hi) >> 1; if (top_sorted_rank[mid] < rk) lo = mid + 1; else hi = mid; }
    if (lo < cnt && top_sorted_rank[lo] == rk) return top_pos[lo];
    return cnt + rk - lo;
}

// pre-order rank among the internal nodes = pre(x) - (leaves before x in pre-order) = pre(x) - first leaf of x
__global__ void k_irank(int n, const int* __restrict__ pre, const int* __restrict__ first, int* __restrict__ irank) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x < n - 1) irank[x] = pre[x] - first[x];
}
// rank-sorted view of the top list (ranks ascending + breadth-first position of each)
__global__ void k_top_ranks(int cnt_max, const int* __restrict__ status, const int* __restrict__ top, const int* __restrict__ irank,
                            int* __restrict__ sorted_rank, int* __restrict__ pos_of) {
    // single block of 1024: bitonic sort of (rank << 11 | bfs position)
    __shared__ long long key[1024];
    const int t = threadIdx.x, cnt = status[2];
    key[t] = (t < cnt) ? (((long long)irank[top[t]] << 11) | t) : 0x7fffffffffffffffLL;
    __syncthreads();
    for (int k = 2; k <= 1024; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            const int ixj = t ^ j;
            if (ixj > t) {
                const bool up = (t & k) == 0;
                const long long a = key[t], b = key[ixj];
                if ((a > b) == up) { key[t] = b; key[ixj] = a; }
            }
            __syncthreads();
        }
    if (t < cnt) { sorted_rank[t] = (int)(key[t] >> 11); pos_of[t] = (int)(key[t] & 2047); }
    (void)cnt_max;
}

__global__ void k_nodes2(int n, const int* __restrict__ left, const int* __restrict__ right, const float* __restrict__ boxes,
                         const int* __restrict__ irank, const int* __restrict__ sorted_rank, const int* __restrict__ pos_of,
                         const int* __restrict__ status, float leaf_guard, TrNode2* __restrict__ out) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n - 1) return;
    const int cnt = status[2], nint = n - 1;
    const int l = left[x], r = right[x];
    const int link0 = l >= nint ? -((l - nint) + 1) : node2_index(l, irank, sorted_rank, pos_of, cnt);
    const int link1 = r >= nint ? -((r - nint) + 1) : node2_index(r, irank, sorted_rank, pos_of, cnt);
    // leaf boxes carry the guard band (the reference never tests a leaf's own box; the band keeps the cull strictly weaker
    // than the triangle test), internal boxes are the exact unions the reference tests
    const float g0 = l >= nint ? leaf_guard : 0.0f, g1 = r >= nint ? leaf_guard : 0.0f;
    const float* b0 = boxes + (size_t)l * 6; const float* b1 = boxes + (size_t)r * 6;
    TrNode2 nd;
    nd.a = make_float4(b0[0] - g0, b0[1] - g0, b0[2] - g0, __int_as_float(link0));
    nd.b = make_float4(b0[3] + g0, b0[4] + g0, b0[5] + g0, __int_as_float(link1));
    nd.c = make_float4(b1[0] - g1, b1[1] - g1, b1[2] - g1, 0.0f);
    nd.d = make_float4(b1[3] + g1, b1[4] + g1, b1[5] + g1, 0.0f);
    out[node2_index(x, irank, sorted_rank, pos_of, cnt)] = nd;
}

// 8-way replicated shared-memory image of a small tree: row (4 i + w) = word w of node i, row (4 nint + 3 k + w) = word w of
// leaf k, every row = the 16-byte word eight times (128 B): lane l reads column l & 7 (trace.cuh, TM_REP)
__global__ void k_small_img(int nint, int nleaves, const float4* __restrict__ nodes2, const float4* __restrict__ leaves, float4* __restrict__ img) {
    const int rows = nint * 4 + nleaves * 3;
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * 8) return;
    const int row = t >> 3;
    img[t] = row < nint * 4 ? nodes2[row] : leaves[row - nint * 4];
}

// ------------------------------------------------------------------ reference-layout views
// bvh_node rows (UtilsFunc.py:24-26, flag arithmetic :219-243 -> leaf 7.0, internal 65534.0)
__global__ void k_ref_views(int n, const int* __restrict__ sorted_prim, const int* __restrict__ left, const int* __restrict__ right,
                            const int* __restrict__ parent, const float* __restrict__ boxes, const int* __restrict__ pre,
                            const int* __restrict__ leafcount, float* __restrict__ bvh_node, float* __restrict__ compact) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int nn = 2 * n - 1;
    if (x >= nn) return;
    bool leaf = x >= n - 1;
    const float* b = boxes + (size_t)x * 6;
    float flags = leaf ? 7.0f : 65534.0f;
    float* o = bvh_node + (size_t)x * 11;
    o[0] = flags;
    o[1] = leaf ? -1.0f : (float)left[x];
    o[2] = leaf ? -1.0f : (float)right[x];
    o[3] = (float)parent[x];
    o[4] = leaf ? (float)sorted_prim[x - (n - 1)] : -1.0f;
#pragma unroll
    for (int q = 0; q < 6; ++q) o[5 + q] = b[q];
    int p = pre[x];
    float* c = compact + (size_t)p * 9;
    c[0] = flags;
    c[1] = leaf ? (float)sorted_prim[x - (n - 1)] : (float)(p + 2 * leafcount[left[x]]);
#pragma unroll
    for (int q = 0; q < 6; ++q) c[2 + q] = b[q];
    c[8] = 0.0f;
}

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

extern "C" int tr_bvh_build(tr_ctx* ctx) {
    if (!ctx || !ctx->d_vertex || ctx->np <= 0) return tr_fail(ctx, TR_ERR_INVALID, "tr_bvh_build: no scene uploaded");
    TR_CUDA(ctx, cudaSetDevice(ctx->device));
    const int n = ctx->np, nn = 2 * n - 1;
    cudaStream_t s = ctx->stream;
    int rc;
    if ((rc = tr_stats_resolve(ctx))) return rc;      // an asynchronous render still owns the timing events used below
    if ((rc = tr_realloc(ctx, &ctx->d_morton_unsorted, (size_t)n * 2))) return rc;
    for (int k = 0; k < 2; ++k) { if ((rc = tr_realloc(ctx, &ctx->d_keys[k], (size_t)n))) return rc; if ((rc = tr_realloc(ctx, &ctx->d_vals[k], (size_t)n))) return rc; }
    if ((rc = tr_realloc(ctx, &ctx->d_left, (size_t)nn))) return rc;
    if ((rc = tr_realloc(ctx, &ctx->d_right, (size_t)nn))) return rc;
    if ((rc = tr_realloc(ctx, &ctx->d_parent, (size_t)nn))) return rc;
    if ((rc = tr_realloc(ctx, &ctx->d_boxes, (size_t)nn * 6))) return rc;
    if ((rc = tr_realloc(ctx, &ctx->d_leafcount, (size_t)nn))) return rc;
    if ((rc = tr_realloc(ctx, &ctx->d_flag, (size_t)nn))) return rc;
    if ((rc = tr_realloc(ctx, &ctx->d_pre, (size_t)nn))) return rc;
    if ((rc = tr_realloc(ctx, &ctx->d_nodes, (size_t)nn))) return rc;
    if ((rc = tr_realloc(ctx, &ctx->d_leaves, (size_t)n))) return rc;
    if ((rc = tr_realloc(ctx, &ctx->d_leaf_of_prim, (size_t)n))) return rc;
    if ((rc = tr_realloc(ctx, &ctx->d_first, (size_t)nn))) return rc;
    if ((rc = tr_realloc(ctx, &ctx->d_sneed, (size_t)nn))) return rc;
    if ((rc = tr_realloc(ctx, &ctx->d_irank, (size_t)nn))) return rc;
    if ((rc = tr_realloc(ctx, &ctx->d_top, (size_t)3 * TR_TOP_MAX))) return rc;          // top ids | their ranks sorted | breadth-first position of each
    if ((rc = tr_realloc(ctx, &ctx->d_nodes2, (size_t)(n > 1 ? n - 1 : 1)))) return rc;
    const size_t img_rows = (size_t)(n - 1) * 4 + (size_t)n * 3;
    const bool small = img_rows * 128 <= TR_SMALL_IMG_MAX;
    if (small && (rc = tr_realloc(ctx, &ctx->d_small_img, img_rows * 8))) return rc;
    const int nblocks = cdiv(n, SORT_CHUNK);
    if ((rc = tr_realloc(ctx, &ctx->d_hist, (size_t)RADIX * nblocks))) return rc;

    TR_CUDA(ctx, cudaEventRecord(ctx->ev0, s));
    TR_CUDA(ctx, cudaMemsetAsync(ctx->d_parent, 0xFF, (size_t)nn * 4, s));
    TR_CUDA(ctx, cudaMemsetAsync(ctx->d_left, 0xFF, (size_t)nn * 4, s));
    TR_CUDA(ctx, cudaMemsetAsync(ctx->d_right, 0xFF, (size_t)nn * 4, s));
    TR_CUDA(ctx, cudaMemsetAsync(ctx->d_flag, 0, (size_t)nn * 4, s));
    TR_CUDA(ctx, cudaMemsetAsync(ctx->d_build_status, 0, 16 * 4, s));

    V3 mn = mk3(ctx->bmin[0], ctx->bmin[1], ctx->bmin[2]), mx = mk3(ctx->bmax[0], ctx->bmax[1], ctx->bmax[2]);
    // guard band of the leaf-box cull: 1e-4 of the scene diagonal (Moller-Trumbore's acceptance region
    // exceeds the exact triangle by ~1e-6 of the ray-origin distance)
    const float ext[3] = {ctx->bmax[0] - ctx->bmin[0], ctx->bmax[1] - ctx->bmin[1], ctx->bmax[2] - ctx->bmin[2]};
    float leaf_guard = 1e-4f * sqrtf(ext[0] * ext[0] + ext[1] * ext[1] + ext[2] * ext[2]);
    if (!(leaf_guard > 0.0f) || !(leaf_guard < 1e30f)) leaf_guard = 1e-4f;
    k_morton<<<cdiv(n, 256), 256, 0, s>>>(ctx->d_vertex, ctx->d_prim, ctx->d_shape, n, mn, mx, ctx->d_keys[0], ctx->d_vals[0], ctx->d_morton_unsorted);
    TR_CHECK_LAUNCH(ctx);
    int cur = 0;
    if (n > 1) {
        for (int pass = 0; pass < 4; ++pass) {           // bits 0..31 cover the reference's bits 0..29 (codes < 2^30)
            int shift = pass * 8;
            k_sort_hist<<<nblocks, SORT_THREADS, 0, s>>>(ctx->d_keys[cur], n, shift, nblocks, ctx->d_hist);
            k_sort_scan<<<1, 1024, 0, s>>>(ctx->d_hist, RADIX * nblocks);
            k_sort_scatter<<<nblocks, SORT_THREADS, 0, s>>>(ctx->d_keys[cur], ctx->d_vals[cur], n, shift, nblocks, ctx->d_hist,
                                                           ctx->d_keys[cur ^ 1], ctx->d_vals[cur ^ 1]);
            TR_CHECK_LAUNCH(ctx);
            cur ^= 1;
        }
        k_karras<<<cdiv(n - 1, 256), 256, 0, s>>>(ctx->d_keys[cur], n, ctx->d_left, ctx->d_right, ctx->d_parent, ctx->d_first);
        TR_CHECK_LAUNCH(ctx);
    }
    ctx->sorted_buf = cur;
    k_refit<<<cdiv(n, 256), 256, 0, s>>>(ctx->d_vertex, ctx->d_prim, ctx->d_shape, ctx->d_vals[cur], n, ctx->d_left, ctx->d_right,
                                        ctx->d_parent, ctx->d_boxes, ctx->d_leafcount, ctx->d_flag, ctx->d_build_status, ctx->d_sneed);
    TR_CHECK_LAUNCH(ctx);
    k_flatten<<<cdiv(nn, 256), 256, 0, s>>>(ctx->d_vertex, ctx->d_prim, ctx->d_shape, ctx->d_vals[cur], n, ctx->d_left, ctx->d_right,
                                           ctx->d_parent, ctx->d_boxes, ctx->d_leafcount, ctx->d_pre, ctx->d_nodes, ctx->d_leaves, ctx->d_leaf_of_prim);
    TR_CHECK_LAUNCH(ctx);
    if (n > 1) {
        int* top = ctx->d_top; int* srank = top + TR_TOP_MAX; int* spos = top + 2 * TR_TOP_MAX;
        k_irank<<<cdiv(n - 1, 256), 256, 0, s>>>(n, ctx->d_pre, ctx->d_first, ctx->d_irank);
        k_top_bfs<<<1, 1024, 0, s>>>(n, ctx->d_left, ctx->d_right, top, ctx->d_build_status);
        k_top_ranks<<<1, 1024, 0, s>>>(TR_TOP_MAX, ctx->d_build_status, top, ctx->d_irank, srank, spos);
        k_nodes2<<<cdiv(n - 1, 256), 256, 0, s>>>(n, ctx->d_left, ctx->d_right, ctx->d_boxes, ctx->d_irank, srank, spos, ctx->d_build_status,
                                                 leaf_guard, ctx->d_nodes2);
        TR_CHECK_LAUNCH(ctx);
    }
    if (small) {
        k_small_img<<<cdiv((int)img_rows * 8, 256), 256, 0, s>>>(n - 1, n, (const float4*)ctx->d_nodes2, (const float4*)ctx->d_leaves, ctx->d_small_img);
        TR_CHECK_LAUNCH(ctx);
    } else if (ctx->d_small_img) { cudaFree(ctx->d_small_img); ctx->capacity.erase((void*)ctx->d_small_img); ctx->d_small_img = nullptr; }
    TR_CUDA(ctx, cudaMemcpyAsync(ctx->h_build_status + 8, ctx->d_boxes, 6 * sizeof(float), cudaMemcpyDeviceToHost, s));   // root box (node 0; the single leaf when n == 1)
    TR_CUDA(ctx, cudaEventRecord(ctx->ev1, s));
    int* status = ctx->h_build_status;                   // pinned: a true asynchronous copy, one wait for build + copy
    TR_CUDA(ctx, cudaMemcpyAsync(status, ctx->d_build_status, 8 * sizeof(int), cudaMemcpyDeviceToHost, s));      // [8..13] hold the root box
    TR_CUDA(ctx, cudaStreamSynchronize(s));
    float ms = 0.0f; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); ctx->stats.ms_build = ms;
    if (status[0] != n - 1)
        return tr_fail(ctx, TR_ERR_AABB, "aabb gen error: %d of %d internal nodes refitted", status[0], n - 1);
    ctx->stack_need = n > 1 ? status[1] : 1; ctx->top_count = n > 1 ? status[2] : 0;
    memcpy(ctx->root_box, status + 8, 6 * sizeof(float));
    if (ctx->stack_need > TR_STACK_MAX)
        return tr_fail(ctx, TR_ERR_STACK, "overflow, need larger stack: the tree needs %d traversal-stack entries (limit %d)", ctx->stack_need, TR_STACK_MAX);
    ctx->bvh_ready = true; ctx->shade_ready = false; ctx->fh_ready = false;
    ctx->gen++;
    return TR_OK;
}

extern "C" int tr_morton_download(tr_ctx* ctx, int32_t* morton_unsorted) {
    if (!ctx || !ctx->bvh_ready || !morton_unsorted) return tr_fail(ctx, TR_ERR_INVALID, "tr_morton_download: BVH not built");
    TR_CUDA(ctx, cudaSetDevice(ctx->device));
    TR_CUDA(ctx, cudaMemcpyAsync(morton_unsorted, ctx->d_morton_unsorted, (size_t)ctx->np * 8, cudaMemcpyDeviceToHost, ctx->stream));
    TR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return TR_OK;
}

__global__ void k_interleave(const int* __restrict__ a, const int* __restrict__ b, int n, int* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { out[i * 2] = a[i]; out[i * 2 + 1] = b[i]; }
}

extern "C" int tr_bvh_download(tr_ctx* ctx, int32_t* morton_sorted, float* bvh_node, float* compact_node) {
    if (!ctx || !ctx->bvh_ready) return tr_fail(ctx, TR_ERR_INVALID, "tr_bvh_download: BVH not built");
    TR_CUDA(ctx, cudaSetDevice(ctx->device));
    const int n = ctx->np, nn = 2 * n - 1;
    cudaStream_t s = ctx->stream;
    if (morton_sorted) {
        int* tmp = nullptr; TR_CUDA(ctx, cudaMalloc((void**)&tmp, (size_t)n * 8));
        k_interleave<<<cdiv(n, 256), 256, 0, s>>>(ctx->d_keys[ctx->sorted_buf], ctx->d_vals[ctx->sorted_buf], n, tmp);
        cudaMemcpyAsync(morton_sorted, tmp, (size_t)n * 8, cudaMemcpyDeviceToHost, s);
        cudaStreamSynchronize(s); cudaFree(tmp);
        TR_CHECK_LAUNCH(ctx);
    }
    if (bvh_node || compact_node) {
        float *d_bn = nullptr, *d_cn = nullptr;
        TR_CUDA(ctx, cudaMalloc((void**)&d_bn, (size_t)nn * 11 * 4));
        TR_CUDA(ctx, cudaMalloc((void**)&d_cn, (size_t)nn * 9 * 4));
        k_ref_views<<<cdiv(nn, 256), 256, 0, s>>>(n, ctx->d_vals[ctx->sorted_buf], ctx->d_left, ctx->d_right, ctx->d_parent,
                                                 ctx->d_boxes, ctx->d_pre, ctx->d_leafcount, d_bn, d_cn);
        if (bvh_node) cudaMemcpyAsync(bvh_node, d_bn, (size_t)nn * 11 * 4, cudaMemcpyDeviceToHost, s);
        if (compact_node) cudaMemcpyAsync(compact_node, d_cn, (size_t)nn * 9 * 4, cudaMemcpyDeviceToHost, s);
        cudaStreamSynchronize(s); cudaFree(d_bn); cudaFree(d_cn);
        TR_CHECK_LAUNCH(ctx);
    }
    return TR_OK;
}
