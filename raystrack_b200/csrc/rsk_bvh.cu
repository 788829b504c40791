// rsk_bvh.cu -- GPU BVH builder: Morton-code LBVH (Karras 2012) collapsed into 96-byte 8-wide quantised nodes.
//
// Replaces utils/bvh.py:14-72 (`build_bvh`, a recursive median split in Python: 20 s for 1M triangles) and the
// permutation step of utils/prepared.py:223-228.  The tree is a different one (closest hits do not depend on
// the tree except at exact t ties); what is kept is the contract: every triangle of the scene is
// reachable, in a traversal-order triangle array, with its mesh id.
//
// Pipeline (all on the device, one stream):
//   1. triangle boxes + centroids, scene bounds (block reduce + ordered-int atomics)
//   2. 63-bit Morton codes, cub radix sort of (code, triangle) pairs
//   3. Karras' binary radix tree over the sorted codes (one thread per internal node)
//   4. bottom-up box refit with per-node arrival counters
//   5. level-by-level collapse: a wide node adopts the binary subtree roots obtained by repeatedly opening the
//      child with the largest surface area until 8 children (sub-trees of <= 3 triangles become leaf children),
//      assigns children to octant slots, quantises their boxes conservatively to 8 bits and emits its triangles
//   6. gather of the triangle / normal records into traversal order
#include <cub/cub.cuh>

#include "rsk_common.cuh"

// Binary tree stage: 0 = Karras radix tree (LBVH, default), 1 = PLOC agglomerative clustering.  On the urban
// benchmark scene both give the same traversal speed (2.49 vs 2.48 Grays/s) and PLOC builds 4x slower (41 vs 9 ms),
// so the radix tree ships; PLOC stays selectable for irregular scenes (profiles/kernel_variants_r1.md).
#ifndef RSK_PLOC
#define RSK_PLOC 0
#endif

#ifndef RSK_BOTTOM_MAX
#define RSK_BOTTOM_MAX 3       // triangles a bottom-level wide node may hold (<= 8 leaf children x 3)
#endif

namespace {

struct Box {
    float3 lo, hi;
};

__device__ __forceinline__ unsigned ord_from_float(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float float_from_ord(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// ---- 1. boxes, centroids, scene bounds.  bounds[0..2] = min (ordered uint), bounds[3..5] = max
__global__ void k_tri_boxes(const float4 *__restrict__ tri, int n, float4 *blo, float4 *bhi, unsigned *bounds) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float3 lo = make_float3(3e38f, 3e38f, 3e38f), hi = make_float3(-3e38f, -3e38f, -3e38f);
    if (i < n) {
        const float4 a = tri[3 * (int64_t)i], e1 = tri[3 * (int64_t)i + 1], e2 = tri[3 * (int64_t)i + 2];
        const float3 p1 = make_float3(a.x + e1.x, a.y + e1.y, a.z + e1.z);
        const float3 p2 = make_float3(a.x + e2.x, a.y + e2.y, a.z + e2.z);
        lo = make_float3(fminf(a.x, fminf(p1.x, p2.x)), fminf(a.y, fminf(p1.y, p2.y)), fminf(a.z, fminf(p1.z, p2.z)));
        hi = make_float3(fmaxf(a.x, fmaxf(p1.x, p2.x)), fmaxf(a.y, fmaxf(p1.y, p2.y)), fmaxf(a.z, fmaxf(p1.z, p2.z)));
        blo[i] = make_float4(lo.x, lo.y, lo.z, a.w);     // .w carries the mesh id (min / max of the sub-tree later)
        bhi[i] = make_float4(hi.x, hi.y, hi.z, a.w);
    }
    typedef cub::BlockReduce<float, 256> BR;
    __shared__ typename BR::TempStorage tmp;
    float v[6] = {lo.x, lo.y, lo.z, hi.x, hi.y, hi.z};
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        float r = c < 3 ? BR(tmp).Reduce(v[c], cub::Min()) : BR(tmp).Reduce(v[c], cub::Max());
        __syncthreads();
        if (threadIdx.x == 0) {
            if (c < 3) atomicMin(&bounds[c], ord_from_float(r));
            else atomicMax(&bounds[c], ord_from_float(r));
        }
    }
}

__device__ __forceinline__ unsigned long long spread21(unsigned long long x) {
    x &= 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

// ---- 2. Morton codes of the box centres
__global__ void k_morton(const float4 *blo, const float4 *bhi, int n, const unsigned *bounds,
                         unsigned long long *codes, unsigned *ids) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float3 smin = make_float3(float_from_ord(bounds[0]), float_from_ord(bounds[1]), float_from_ord(bounds[2]));
    const float3 smax = make_float3(float_from_ord(bounds[3]), float_from_ord(bounds[4]), float_from_ord(bounds[5]));
    const float4 lo = blo[i], hi = bhi[i];
#ifndef RSK_MORTON_UNIFORM
#define RSK_MORTON_UNIFORM 1
#endif
#if RSK_MORTON_UNIFORM
    // one scale for the three axes: Morton cells stay cubic even when the scene is flat (a city is 7x wider than tall)
    const float ext = fmaxf(fmaxf(smax.x - smin.x, smax.y - smin.y), fmaxf(smax.z - smin.z, 1e-30f));
    const float ex = ext, ey = ext, ez = ext;
#else
    const float ex = fmaxf(smax.x - smin.x, 1e-30f), ey = fmaxf(smax.y - smin.y, 1e-30f), ez = fmaxf(smax.z - smin.z, 1e-30f);
#endif
    const float cx = (0.5f * (lo.x + hi.x) - smin.x) / ex, cy = (0.5f * (lo.y + hi.y) - smin.y) / ey, cz = (0.5f * (lo.z + hi.z) - smin.z) / ez;
    const float s = 2097151.0f;
    const unsigned long long qx = (unsigned long long)fminf(fmaxf(cx * s, 0.f), s);
    const unsigned long long qy = (unsigned long long)fminf(fmaxf(cy * s, 0.f), s);
    const unsigned long long qz = (unsigned long long)fminf(fmaxf(cz * s, 0.f), s);
    codes[i] = spread21(qx) | (spread21(qy) << 1) | (spread21(qz) << 2);
    ids[i] = (unsigned)i;
}

// ---- 3. Karras' radix tree.  Node ids: internal i in [0,n-1); leaf j (sorted position) = (n-1)+j.
__device__ __forceinline__ int delta(const unsigned long long *codes, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const unsigned long long x = codes[i] ^ codes[j];
    return x == 0ull ? 64 + __clz(i ^ j) : __clzll((long long)x);
}

__attribute__((unused)) __global__ void k_radix_tree(const unsigned long long *codes, int n, int *left, int *right, int *parent, int *count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = delta(codes, n, i, i + 1) - delta(codes, n, i, i - 1) >= 0 ? 1 : -1;
    const int dmin = delta(codes, n, i, i - d);
    int lmax = 2;
    while (delta(codes, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(codes, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta(codes, n, i, j);
    int s = 0, t = l;
    do {
        t = (t + 1) >> 1;
        if (delta(codes, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const int lc = (lo == gamma) ? (n - 1) + gamma : gamma;
    const int rc = (hi == gamma + 1) ? (n - 1) + gamma + 1 : gamma + 1;
    left[i] = lc;
    right[i] = rc;
    parent[lc] = i;
    parent[rc] = i;
    count[i] = hi - lo + 1;
    if (i == 0) parent[0] = -1;
}


// ---- 3b. PLOC (parallel locally-ordered clustering, Meister & Bittner 2018): bottom-up agglomeration over the
// Morton-ordered cluster array.  Every round each cluster picks the neighbour within RSK_PLOC_RADIUS positions
// whose union box has the smallest surface area; mutual choices merge; the array is compacted, order preserved.
// Produces markedly better trees than the radix tree (fewer node visits per ray) for a few ms more build time.
#ifndef RSK_PLOC_RADIUS
#define RSK_PLOC_RADIUS 16
#endif
__global__ void k_ploc_init(const unsigned *ids, const float4 *tlo, const float4 *thi, int n, float4 *nlo, float4 *nhi, int *cluster) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int node = (n - 1) + j;
    nlo[node] = tlo[ids[j]];
    nhi[node] = thi[ids[j]];
    cluster[j] = node;
}

__global__ void k_ploc_nearest(const int *cluster, int nc, const float4 *nlo, const float4 *nhi, int *nearest) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nc) return;
    const float4 lo = nlo[cluster[i]], hi = nhi[cluster[i]];
    float best = 3e38f;
    int bj = -1;
    const int j0 = max(0, i - RSK_PLOC_RADIUS), j1 = min(nc - 1, i + RSK_PLOC_RADIUS);
    // the pair partner i^1 is examined first and wins ties, so that degenerate inputs (identical boxes) still pair
    // up (0,1),(2,3),... and the number of clusters halves every round
    for (int t = -1; t <= j1 - j0; ++t) {
        const int j = t < 0 ? (i ^ 1) : j0 + t;
        if (j == i || j >= nc || (t >= 0 && j == (i ^ 1))) continue;
        const float4 l2 = nlo[cluster[j]], h2 = nhi[cluster[j]];
        const float x = fmaxf(hi.x, h2.x) - fminf(lo.x, l2.x), y = fmaxf(hi.y, h2.y) - fminf(lo.y, l2.y), z = fmaxf(hi.z, h2.z) - fminf(lo.z, l2.z);
        const float ar = x * y + y * z + z * x;
        if (ar < best) { best = ar; bj = j; }
    }
    nearest[i] = bj;
}

__global__ void k_ploc_merge(const int *cluster, const int *nearest, int nc, int n, float4 *nlo, float4 *nhi, int *left, int *right,
                             int *count, int *node_counter, int *merged, int *keep) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nc) return;
    const int j = nearest[i];
    int out = cluster[i], k = 1;
    if (j >= 0 && nearest[j] == i) {
        if (i < j) {
            const int a = cluster[i], b = cluster[j];
            const int p = atomicAdd(node_counter, 1);
            left[p] = a;
            right[p] = b;
            count[p] = (a >= n - 1 ? 1 : count[a]) + (b >= n - 1 ? 1 : count[b]);
            const float4 la = nlo[a], lb = nlo[b], ha = nhi[a], hb = nhi[b];
            nlo[p] = make_float4(fminf(la.x, lb.x), fminf(la.y, lb.y), fminf(la.z, lb.z), __int_as_float(min(__float_as_int(la.w), __float_as_int(lb.w))));
            nhi[p] = make_float4(fmaxf(ha.x, hb.x), fmaxf(ha.y, hb.y), fmaxf(ha.z, hb.z), __int_as_float(max(__float_as_int(ha.w), __float_as_int(hb.w))));
            out = p;
        } else {
            k = 0;
        }
    }
    merged[i] = out;
    keep[i] = k;
}

__global__ void k_ploc_compact(const int *merged, const int *keep, const int *pos, int nc, int *cluster_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nc) return;
    if (keep[i]) cluster_out[pos[i]] = merged[i];
}

// ---- 4. refit: leaf boxes from the sorted triangles, internal boxes bottom-up
__global__ void k_refit(const unsigned *ids, const float4 *tlo, const float4 *thi, int n, const int *left, const int *right,
                        const int *parent, float4 *nlo, float4 *nhi, int *arrivals) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    int node = (n - 1) + j;
    nlo[node] = tlo[ids[j]];
    nhi[node] = thi[ids[j]];
    __threadfence();
    int p = parent[node];
    while (p >= 0) {
        if (atomicAdd(&arrivals[p], 1) == 0) return;     // the sibling finishes this node
        __threadfence();
        const int l = left[p], r = right[p];
        const float4 a = __ldcg(nlo + l), b = __ldcg(nlo + r), c = __ldcg(nhi + l), e = __ldcg(nhi + r);
        nlo[p] = make_float4(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z),
                             __int_as_float(min(__float_as_int(a.w), __float_as_int(b.w))));
        nhi[p] = make_float4(fmaxf(c.x, e.x), fmaxf(c.y, e.y), fmaxf(c.z, e.z),
                             __int_as_float(max(__float_as_int(c.w), __float_as_int(e.w))));
        __threadfence();
        p = parent[p];
    }
}

// Conservative padding (world units) and the lower bound of the (unbiased) quantisation exponent, from the largest
// coordinate of the scene: computed by one thread on the device so that the build never waits for the bounds.
struct BuildParams {
    float pad;
    int min_exp;
};
__global__ void k_build_params(const unsigned *bounds, BuildParams *bp) {
    float max_abs = 0.f;
    for (int c = 0; c < 6; ++c) max_abs = fmaxf(max_abs, fabsf(float_from_ord(bounds[c])));
    if (!(max_abs > 0.f)) max_abs = 1.f;
    bp->pad = max_abs * 4.76837158e-7f;                 // 2^-21 of the largest coordinate: several ulps
    int me;
    frexpf(max_abs, &me);
    bp->min_exp = me - 20;                              // grid step never below ~2^-20 of the coordinates
}

// ---- 5. collapse one level of wide nodes
struct CollapseArgs {
    const int *left, *right;
    const int *count;             // triangles below each binary node (internal nodes only; leaves count 1)
    int root;                     // binary root id
    const float4 *nlo, *nhi;
    const unsigned *ids;          // sorted position -> input triangle
    int n;                        // triangles
    const int2 *queue_in;         // (binary node, wide index)
    const int *n_in;              // entries of queue_in: written by the previous level's launch, never read by the host
    int2 *queue_out;
    int *n_out;
    int *depth;                   // deepest level that held a node (atomicMax)
    int level;
    int *node_counter;            // next free wide-node index
    int *tri_counter;             // next free traversal-order triangle slot
    WideNode *nodes;
    int *tri_order;               // traversal slot -> input triangle
    const BuildParams *bp;        // padding and exponent floor, derived from the scene bounds on the device
};

__device__ __forceinline__ int sub_count(const CollapseArgs &a, int node) {
    return node >= a.n - 1 ? 1 : a.count[node];
}
// input triangles of a sub-tree of at most RSK_LEAF_MAX leaves, left to right
__device__ __forceinline__ int sub_triangles(const CollapseArgs &a, int node, int *out) {
    int stack[RSK_LEAF_MAX + 1], sp = 0, m = 0;
    stack[sp++] = node;
    while (sp > 0) {
        const int v = stack[--sp];
        if (v >= a.n - 1) out[m++] = (int)a.ids[v - (a.n - 1)];
        else { stack[sp++] = a.right[v]; stack[sp++] = a.left[v]; }
    }
    return m;
}
__device__ __forceinline__ float half_area(const float4 &lo, const float4 &hi) {
    const float x = hi.x - lo.x, y = hi.y - lo.y, z = hi.z - lo.z;
    return x * y + y * z + z * x;
}

__device__ __forceinline__ int quant_exp(float ext, int min_exp) {
    int e = min_exp;
    if (ext > 0.f) {
        int k;
        frexpf(ext / 255.0f, &k);     // ext/255 = m * 2^k, m in [0.5,1)  =>  2^k >= ext/255
        e = max(k, min_exp);
    }
    return min(max(e, -126), 112);      // + 14 on RSK_PRMT_AXES axes must stay a finite float
}

// One wide node of the level: grid-stride over the level's queue, whose length only the device knows (the host
// launches a fixed sequence of levels; a launch past the last level finds an empty queue and returns).
__device__ void collapse_one(const CollapseArgs &a, int q, float a_pad, int a_min_exp);

__global__ void k_collapse(const CollapseArgs a) {
    const int n_in = *a.n_in;
    if (n_in <= 0) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicMax(a.depth, a.level + 1);
    const float pad = a.bp->pad;
    const int min_exp = a.bp->min_exp;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n_in; q += gridDim.x * blockDim.x) collapse_one(a, q, pad, min_exp);
}

__device__ void collapse_one(const CollapseArgs &a, int q, float a_pad, int a_min_exp) {
    const int bnode = a.queue_in[q].x, widx = a.queue_in[q].y;

    // Which binary sub-trees become the (up to) 8 children.  Sub-trees of <= RSK_LEAF_MAX triangles are leaf
    // children.  A node over <= RSK_BOTTOM_MAX triangles is a *bottom* node: it keeps opening its largest child
    // (by triangle count) so that everything ends up in leaf children.  Above that, only children too big to
    // become a bottom node are opened (largest surface area first); a child of 4..RSK_BOTTOM_MAX triangles is kept
    // whole and becomes one well-filled bottom node instead of being shredded into several 2-child nodes.
    int cand[RSK_WIDE];
    int nc = 2;
    cand[0] = a.left[bnode];
    cand[1] = a.right[bnode];
    const bool bottom = sub_count(a, bnode) <= RSK_BOTTOM_MAX;
    while (nc < RSK_WIDE) {
        int best = -1;
        float best_score = -1.f;
        for (int c = 0; c < nc; ++c) {
            const int cnt = sub_count(a, cand[c]);
            if (cnt <= (bottom ? RSK_LEAF_MAX : RSK_BOTTOM_MAX)) continue;
            const float score = bottom ? (float)cnt : half_area(a.nlo[cand[c]], a.nhi[cand[c]]);
            if (score > best_score) { best_score = score; best = c; }
        }
        if (best < 0) break;
        const int open = cand[best];
        cand[best] = a.left[open];
        cand[nc++] = a.right[open];
    }

    // padded child boxes, node box
    float3 clo[RSK_WIDE], chi[RSK_WIDE];
    float3 lo = make_float3(3e38f, 3e38f, 3e38f), hi = make_float3(-3e38f, -3e38f, -3e38f);
    for (int c = 0; c < nc; ++c) {
        const float4 l = a.nlo[cand[c]], h = a.nhi[cand[c]];
        clo[c] = make_float3(l.x - a_pad, l.y - a_pad, l.z - a_pad);
        chi[c] = make_float3(h.x + a_pad, h.y + a_pad, h.z + a_pad);
        lo = make_float3(fminf(lo.x, clo[c].x), fminf(lo.y, clo[c].y), fminf(lo.z, clo[c].z));
        hi = make_float3(fmaxf(hi.x, chi[c].x), fmaxf(hi.y, chi[c].y), fmaxf(hi.z, chi[c].z));
    }
    const float3 ctr = make_float3(0.5f * (lo.x + hi.x), 0.5f * (lo.y + hi.y), 0.5f * (lo.z + hi.z));

    // octant slots: greedy maximum of  (child centre - node centre) . (+-1,+-1,+-1)
    int slot_of[RSK_WIDE], child_in[RSK_WIDE];
    for (int s = 0; s < RSK_WIDE; ++s) { slot_of[s] = -1; child_in[s] = -1; }
    for (int round = 0; round < nc; ++round) {
        float bestv = -3e38f;
        int bc = -1, bs = -1;
        for (int c = 0; c < nc; ++c) {
            if (slot_of[c] >= 0) continue;
            const float vx = 0.5f * (clo[c].x + chi[c].x) - ctr.x, vy = 0.5f * (clo[c].y + chi[c].y) - ctr.y, vz = 0.5f * (clo[c].z + chi[c].z) - ctr.z;
            for (int s = 0; s < RSK_WIDE; ++s) {
                if (child_in[s] >= 0) continue;
                const float v = ((s & 1) ? vx : -vx) + ((s & 2) ? vy : -vy) + ((s & 4) ? vz : -vz);
                if (v > bestv) { bestv = v; bc = c; bs = s; }
            }
        }
        slot_of[bc] = bs;
        child_in[bs] = bc;
    }

    int n_inner = 0, n_leaf_tris = 0;
    for (int s = 0; s < RSK_WIDE; ++s) {
        if (child_in[s] < 0) continue;
        const int cnt = sub_count(a, cand[child_in[s]]);
        if (cnt <= RSK_LEAF_MAX) n_leaf_tris += cnt; else n_inner++;
    }
    const int child_base = n_inner ? atomicAdd(a.node_counter, n_inner) : 0;
    const int tri_base = n_leaf_tris ? atomicAdd(a.tri_counter, n_leaf_tris) : 0;
    int out_base = n_inner ? atomicAdd(a.n_out, n_inner) : 0;

    WideNode node;
    node.sid_min = __float_as_int(a.nlo[bnode].w);
    node.sid_max = __float_as_int(a.nhi[bnode].w);
    node.reserved = 0u;
    node.ox = lo.x; node.oy = lo.y; node.oz = lo.z;
    const int ex = quant_exp(hi.x - lo.x, a_min_exp), ey = quant_exp(hi.y - lo.y, a_min_exp), ez = quant_exp(hi.z - lo.z, a_min_exp);
    node.scale[0] = exp2f((float)(ex + ((RSK_PRMT_AXES & 1) ? 14 : 0)));
    node.scale[1] = exp2f((float)(ey + ((RSK_PRMT_AXES & 2) ? 14 : 0)));
    node.scale[2] = exp2f((float)(ez + ((RSK_PRMT_AXES & 4) ? 14 : 0)));
    const float ix = exp2f((float)-ex), iy = exp2f((float)-ey), iz = exp2f((float)-ez);
    node.child_base = (uint32_t)child_base;
    node.tri_base = (uint32_t)tri_base;
    uint32_t imask = 0u, leaf_bits = 0u;
    int inner_rank = 0, tri_off = 0;
    for (int s = 0; s < RSK_WIDE; ++s) {
        const int c = child_in[s];
        if (c < 0) {        // empty slot: an inverted box no ray can enter, no bits
            for (int ax = 0; ax < 3; ++ax) { node.qlo[ax][s] = 255; node.qhi[ax][s] = 0; }
            continue;
        }
        node.qlo[0][s] = (uint8_t)fminf(fmaxf(floorf((clo[c].x - lo.x) * ix), 0.f), 255.f);
        node.qlo[1][s] = (uint8_t)fminf(fmaxf(floorf((clo[c].y - lo.y) * iy), 0.f), 255.f);
        node.qlo[2][s] = (uint8_t)fminf(fmaxf(floorf((clo[c].z - lo.z) * iz), 0.f), 255.f);
        node.qhi[0][s] = (uint8_t)fminf(fmaxf(ceilf((chi[c].x - lo.x) * ix), 0.f), 255.f);
        node.qhi[1][s] = (uint8_t)fminf(fmaxf(ceilf((chi[c].y - lo.y) * iy), 0.f), 255.f);
        node.qhi[2][s] = (uint8_t)fminf(fmaxf(ceilf((chi[c].z - lo.z) * iz), 0.f), 255.f);
        const int bn = cand[c];
        const int cnt = sub_count(a, bn);
        if (cnt <= RSK_LEAF_MAX) {
            int tris[RSK_LEAF_MAX];
            sub_triangles(a, bn, tris);
            for (int t = 0; t < cnt; ++t) a.tri_order[tri_base + tri_off + t] = tris[t];
            leaf_bits |= ((1u << cnt) - 1u) << (3 * s);
            tri_off += cnt;
        } else {
            imask |= 1u << s;
            a.queue_out[out_base + inner_rank] = make_int2(bn, child_base + inner_rank);
            inner_rank++;
        }
    }
    node.leaf_imask = leaf_bits | (imask << 24);
    a.nodes[widx] = node;
}

// ---- 6. gather triangle records into traversal order
__global__ void k_gather(const float4 *tri_in, const float4 *nrm_in, const int *order, int n, float4 *tri_out, float4 *nrm_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t s = order[i];
    tri_out[3 * (int64_t)i] = tri_in[3 * s];
    tri_out[3 * (int64_t)i + 1] = tri_in[3 * s + 1];
    tri_out[3 * (int64_t)i + 2] = tri_in[3 * s + 2];
    nrm_out[i] = nrm_in[s];
}

__global__ void k_iota(int *p, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}

// tiny scenes (<= RSK_LEAF_MAX triangles): a root whose only child is a leaf with every triangle
__global__ void k_tiny_root(const float4 *tlo, const float4 *thi, int n, WideNode *nodes, const BuildParams *bp) {
    const float pad = bp->pad;
    const int min_exp = bp->min_exp;
    float3 lo = make_float3(3e38f, 3e38f, 3e38f), hi = make_float3(-3e38f, -3e38f, -3e38f);
    int smin = 0x7fffffff, smax = -1;
    for (int i = 0; i < n; ++i) {
        lo = make_float3(fminf(lo.x, tlo[i].x - pad), fminf(lo.y, tlo[i].y - pad), fminf(lo.z, tlo[i].z - pad));
        hi = make_float3(fmaxf(hi.x, thi[i].x + pad), fmaxf(hi.y, thi[i].y + pad), fmaxf(hi.z, thi[i].z + pad));
        smin = min(smin, __float_as_int(tlo[i].w));
        smax = max(smax, __float_as_int(tlo[i].w));
    }
    WideNode node;
    memset(&node, 0, sizeof(node));
    node.sid_min = smin;
    node.sid_max = smax;
    node.ox = lo.x; node.oy = lo.y; node.oz = lo.z;
    const int e[3] = {quant_exp(hi.x - lo.x, min_exp), quant_exp(hi.y - lo.y, min_exp), quant_exp(hi.z - lo.z, min_exp)};
    for (int ax = 0; ax < 3; ++ax) node.scale[ax] = exp2f((float)(e[ax] + (((RSK_PRMT_AXES >> ax) & 1) ? 14 : 0)));
    for (int s = 0; s < RSK_WIDE; ++s)
        for (int ax = 0; ax < 3; ++ax) { node.qlo[ax][s] = 255; node.qhi[ax][s] = 0; }
    for (int ax = 0; ax < 3; ++ax) { node.qlo[ax][0] = 0; node.qhi[ax][0] = 255; }
    node.leaf_imask = (1u << n) - 1u;      // slot 0: a leaf with every triangle
    nodes[0] = node;
}

}  // namespace

int rsk_bvh_build(rsk_scene *sc, const float4 *tri_in, const float4 *nrm_in) {
    rsk_ctx *ctx = sc->ctx;
    cudaStream_t s = ctx->stream;
    const int n = (int)sc->n_tri;
    RSK_REQUIRE(n >= 1, "rsk_bvh_build: empty scene");
    RSK_REQUIRE(sc->n_tri < (1ll << 30), "rsk_bvh_build: too many triangles");

    cudaEvent_t t0, t1;
    RSK_CUDA(cudaEventCreate(&t0));
    RSK_CUDA(cudaEventCreate(&t1));
    RSK_CUDA(cudaEventRecord(t0, s));

    float4 *tlo = nullptr, *thi = nullptr, *nlo = nullptr, *nhi = nullptr;
    unsigned *bounds = nullptr, *ids = nullptr, *ids_sorted = nullptr;
    unsigned long long *codes = nullptr, *codes_sorted = nullptr;
    int *left = nullptr, *right = nullptr, *parent = nullptr, *count = nullptr, *arrivals = nullptr;
    int *pl_cluster[2] = {nullptr, nullptr}, *pl_nearest = nullptr, *pl_merged = nullptr, *pl_keep = nullptr, *pl_pos = nullptr;
    void *scan_tmp = nullptr;
    int2 *queue[2] = {nullptr, nullptr};
    int *counters = nullptr;      // [0] node counter, [1] tri counter, [2] depth, [4 + L] queue length of level L
    BuildParams *bparams = nullptr;
    void *sort_tmp = nullptr;
    WideNode *nodes = nullptr;
    int rc = RSK_OK;
    auto cleanup = [&]() {
        rsk_dev_free(tlo); rsk_dev_free(thi); rsk_dev_free(nlo); rsk_dev_free(nhi); rsk_dev_free(bounds); rsk_dev_free(ids); rsk_dev_free(ids_sorted);
        rsk_dev_free(codes); rsk_dev_free(codes_sorted); rsk_dev_free(left); rsk_dev_free(right); rsk_dev_free(parent); rsk_dev_free(count);
        rsk_dev_free(pl_cluster[0]); rsk_dev_free(pl_cluster[1]); rsk_dev_free(pl_nearest); rsk_dev_free(pl_merged); rsk_dev_free(pl_keep); rsk_dev_free(pl_pos); rsk_dev_free(scan_tmp);
        rsk_dev_free(arrivals); rsk_dev_free(queue[0]); rsk_dev_free(queue[1]); rsk_dev_free(counters); rsk_dev_free(sort_tmp); rsk_dev_free(bparams);
        cudaEventDestroy(t0); cudaEventDestroy(t1);
    };
#define B_TRY(expr) do { rc = (expr); if (rc != RSK_OK) { cleanup(); rsk_dev_free(nodes); return rc; } } while (0)
#define B_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { rsk_set_error("%s failed: %s", #call, cudaGetErrorString(e__)); cleanup(); rsk_dev_free(nodes); return RSK_ERR_CUDA; } } while (0)

    B_TRY(rsk_dev_alloc(&tlo, n)); B_TRY(rsk_dev_alloc(&thi, n));
    B_TRY(rsk_dev_alloc(&bounds, 6));
    {
        const unsigned init[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
        B_CUDA(cudaMemcpyAsync(bounds, init, sizeof(init), cudaMemcpyHostToDevice, s));
    }
    k_tri_boxes<<<rsk_blocks(n, 256), 256, 0, s>>>(tri_in, n, tlo, thi, bounds);
    ctx->launches++;
    // padding / exponent floor from the scene bounds, on the device (no read-back: the build waits for the GPU once, at its end)
    B_TRY(rsk_dev_alloc(&bparams, 1));
    k_build_params<<<1, 1, 0, s>>>(bounds, bparams);
    ctx->launches++;

    B_TRY(rsk_dev_alloc(&sc->tri_index, n));
    const int max_nodes = n > RSK_LEAF_MAX ? n : 1;
    B_TRY(rsk_dev_alloc(&nodes, max_nodes));

    int depth = 1, n_nodes = 1;
    if (n <= RSK_LEAF_MAX) {
        k_tiny_root<<<1, 1, 0, s>>>(tlo, thi, n, nodes, bparams);
        k_iota<<<1, 32, 0, s>>>(sc->tri_index, n);
        ctx->launches += 2;
    } else {
        B_TRY(rsk_dev_alloc(&codes, n)); B_TRY(rsk_dev_alloc(&codes_sorted, n));
        B_TRY(rsk_dev_alloc(&ids, n)); B_TRY(rsk_dev_alloc(&ids_sorted, n));
        k_morton<<<rsk_blocks(n, 256), 256, 0, s>>>(tlo, thi, n, bounds, codes, ids);
        ctx->launches++;
        size_t tmp_bytes = 0;
        B_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, codes, codes_sorted, ids, ids_sorted, n, 0, 63, s));
        { unsigned char *tmp = nullptr; B_TRY(rsk_dev_alloc(&tmp, tmp_bytes)); sort_tmp = tmp; }
        B_CUDA(cub::DeviceRadixSort::SortPairs(sort_tmp, tmp_bytes, codes, codes_sorted, ids, ids_sorted, n, 0, 63, s));
        ctx->launches += 8;

        B_TRY(rsk_dev_alloc(&left, n - 1)); B_TRY(rsk_dev_alloc(&right, n - 1)); B_TRY(rsk_dev_alloc(&count, n - 1));
        B_TRY(rsk_dev_alloc(&nlo, 2 * (size_t)n - 1)); B_TRY(rsk_dev_alloc(&nhi, 2 * (size_t)n - 1));
        int root_id = 0;
#if RSK_PLOC
        {
            B_TRY(rsk_dev_alloc(&pl_cluster[0], n)); B_TRY(rsk_dev_alloc(&pl_cluster[1], n));
            B_TRY(rsk_dev_alloc(&pl_nearest, n)); B_TRY(rsk_dev_alloc(&pl_merged, n)); B_TRY(rsk_dev_alloc(&pl_keep, n));
            B_TRY(rsk_dev_alloc(&pl_pos, n)); B_TRY(rsk_dev_alloc(&counters, 4));
            size_t scan_bytes = 0;
            B_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, pl_keep, pl_pos, n, s));
            { unsigned char *tmp = nullptr; B_TRY(rsk_dev_alloc(&tmp, scan_bytes)); scan_tmp = tmp; }
            B_CUDA(cudaMemsetAsync(counters, 0, 4 * sizeof(int), s));
            k_ploc_init<<<rsk_blocks(n, 256), 256, 0, s>>>(ids_sorted, tlo, thi, n, nlo, nhi, pl_cluster[0]);
            ctx->launches++;
            int nc = n, cur_c = 0, rounds = 0;
            while (nc > 1) {
                k_ploc_nearest<<<rsk_blocks(nc, 128), 128, 0, s>>>(pl_cluster[cur_c], nc, nlo, nhi, pl_nearest);
                k_ploc_merge<<<rsk_blocks(nc, 256), 256, 0, s>>>(pl_cluster[cur_c], pl_nearest, nc, n, nlo, nhi, left, right, count, counters,
                                                                 pl_merged, pl_keep);
                B_CUDA(cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, pl_keep, pl_pos, nc, s));
                k_ploc_compact<<<rsk_blocks(nc, 256), 256, 0, s>>>(pl_merged, pl_keep, pl_pos, nc, pl_cluster[cur_c ^ 1]);
                ctx->launches += 4;
                int h_nodes = 0;
                B_CUDA(cudaMemcpyAsync(&h_nodes, counters, sizeof(int), cudaMemcpyDeviceToHost, s));
                B_CUDA(cudaStreamSynchronize(s));
                nc = n - h_nodes;                          // every merge removes one cluster
                cur_c ^= 1;
                if (++rounds > 4096) { rsk_set_error("rsk_bvh_build: clustering did not converge"); cleanup(); rsk_dev_free(nodes); return RSK_ERR_CUDA; }
            }
            B_CUDA(cudaMemcpyAsync(&root_id, pl_cluster[cur_c], sizeof(int), cudaMemcpyDeviceToHost, s));
            B_CUDA(cudaStreamSynchronize(s));
            rsk_dev_free(counters); counters = nullptr;
        }
#else
        B_TRY(rsk_dev_alloc(&parent, 2 * (size_t)n - 1)); B_TRY(rsk_dev_alloc(&arrivals, n - 1));
        B_CUDA(cudaMemsetAsync(arrivals, 0, (size_t)(n - 1) * sizeof(int), s));
        k_radix_tree<<<rsk_blocks(n - 1, 256), 256, 0, s>>>(codes_sorted, n, left, right, parent, count);
        k_refit<<<rsk_blocks(n, 256), 256, 0, s>>>(ids_sorted, tlo, thi, n, left, right, parent, nlo, nhi, arrivals);
        ctx->launches += 2;
#endif

        // Level-by-level collapse without host round trips: level L reads its queue length from counters[4 + L] and
        // appends to counters[5 + L]; the host launches a fixed run of levels (grids sized by the bound min(8^L, n),
        // grid-stride inside) and reads the counters back once.  Trees deeper than the first run (rare: the collapse
        // opens the largest child first, 1M triangles give 9-10 levels) get further runs of four levels.
        B_TRY(rsk_dev_alloc(&queue[0], n)); B_TRY(rsk_dev_alloc(&queue[1], n));
        constexpr int LEVEL_SLOTS = 4 * RSK_MAX_DEPTH_HOST + 8;
        B_TRY(rsk_dev_alloc(&counters, 4 + LEVEL_SLOTS));
        {
            int init_counters[4 + LEVEL_SLOTS];
            memset(init_counters, 0, sizeof(init_counters));
            init_counters[0] = 1;                              // wide node 0 = the root
            init_counters[4] = 1;                              // level 0 holds the root
            B_CUDA(cudaMemcpyAsync(counters, init_counters, sizeof(init_counters), cudaMemcpyHostToDevice, s));
        }
        const int2 root = make_int2(root_id, 0);
        B_CUDA(cudaMemcpyAsync(queue[0], &root, sizeof(root), cudaMemcpyHostToDevice, s));
        int level = 0, cur = 0, pending = 1;
        int run = 12;
        while (pending > 0 && level < 4 * RSK_MAX_DEPTH_HOST) {
            for (int k = 0; k < run; ++k, ++level) {
                CollapseArgs a;
                a.left = left; a.right = right; a.count = count; a.root = root_id; a.nlo = nlo; a.nhi = nhi; a.ids = ids_sorted; a.n = n;
                a.queue_in = queue[cur]; a.n_in = counters + 4 + level; a.queue_out = queue[cur ^ 1]; a.n_out = counters + 5 + level;
                a.depth = counters + 2; a.level = level;
                a.node_counter = counters; a.tri_counter = counters + 1; a.nodes = nodes; a.tri_order = sc->tri_index;
                a.bp = bparams;
                double bound = 1.0;
                for (int l = 0; l < level && bound < (double)n; ++l) bound *= RSK_WIDE;
                const int64_t width = (int64_t)(bound < (double)n ? bound : (double)n);
                const int64_t blocks = rsk_blocks(width, 128);
                k_collapse<<<(unsigned)(blocks < 4096 ? blocks : 4096), 128, 0, s>>>(a);
                ctx->launches++;
                cur ^= 1;
            }
            int h[4 + LEVEL_SLOTS];
            B_CUDA(cudaMemcpyAsync(h, counters, sizeof(h), cudaMemcpyDeviceToHost, s));
            B_CUDA(cudaStreamSynchronize(s));
            n_nodes = h[0];
            depth = h[2];
            pending = h[4 + level];
            run = 4;
        }
        if (depth > RSK_MAX_DEPTH_HOST) {
            rsk_set_error("rsk_bvh_build: wide tree depth %d exceeds the traversal stack (%d)", depth, RSK_MAX_DEPTH_HOST);
            cleanup(); rsk_dev_free(nodes);
            return RSK_ERR_INVALID;
        }
    }

    // gather triangles, shrink the node array
    B_TRY(rsk_dev_alloc(&sc->tri, 3 * (size_t)n));
    B_TRY(rsk_dev_alloc(&sc->nrm, n));
    k_gather<<<rsk_blocks(n, 256), 256, 0, s>>>(tri_in, nrm_in, sc->tri_index, n, sc->tri, sc->nrm);
    ctx->launches++;
    uint4 *packed = nullptr;
    B_TRY(rsk_dev_alloc(&packed, (size_t)n_nodes * RSK_NODE_WORDS));
    B_CUDA(cudaMemcpyAsync(packed, nodes, (size_t)n_nodes * sizeof(WideNode), cudaMemcpyDeviceToDevice, s));
    B_CUDA(cudaEventRecord(t1, s));
    B_CUDA(cudaStreamSynchronize(s));
    B_CUDA(cudaGetLastError());
    float ms = 0.f;
    cudaEventElapsedTime(&ms, t0, t1);
    sc->nodes = packed;
    sc->n_nodes = n_nodes;
    sc->depth = depth;
    sc->build_us = (int64_t)(ms * 1000.f);
    rsk_dev_free(nodes);
    cleanup();
#undef B_TRY
#undef B_CUDA
    return RSK_OK;
}
