// cluster.cuh — K6 cluster_{dir,cc,adj} and K7 emit_compact.
//
// Directional / cc (src/algo/directional.rs:30-54, :74-88): the reference visits UMIs by
// frequency descending and, from each UMI still present, removes everything reachable through
// edges u -> v (dist <= k, freq[v] <= thr[u]) by a recursive DFS.  The set removed after a root is
// closed under reachability, so   v is emitted  <=>  no earlier-visited UMI reaches v,   and the
// root that removes v is the earliest-visited UMI that reaches it.  That is a minimum-label
// propagation over the directed edge list: label[u] starts as its visit rank, every sweep does
// label[dst] = min(label[dst], label[src]) until nothing changes; the fixpoint is unique, so the
// result does not depend on thread scheduling (bit-exact, order-dependent clustering without
// replaying the recursion).
//
// Upstream-intended adjacency (opt-in; the reference's adj keeps everything): greedy in
// visit order without recursion = lexicographically-first maximal independent set, resolved in
// rounds: a UMI is kept once all its earlier neighbours are removed, removed once one is kept.
#pragma once
#include "common.cuh"
#include "pack.cuh"

// STAMPED = false: the first sweeps, where nearly every source has just been lowered — reading the stamp would only add a
// gather per edge; lowerings are stamped all the same so that later sweeps can skip.
template <bool STAMPED>
__global__ void __launch_bounds__(256) label_sweep_kernel(const uint2 *__restrict__ edges, u64 n_edges_host, const unsigned long long *n_ptr,
                                                          unsigned long long *label, DevScalars *sc, u32 *stamp, u32 sweep) {
    const u64 n_edges = n_ptr ? *n_ptr : n_edges_host;       // the contracted list's length lives on the device
    u64 stride = (u64)gridDim.x * 256;
    u32 any = 0;
    for (u64 e = (u64)blockIdx.x * 256 + threadIdx.x; e < n_edges; e += stride) {
        uint2 ed = edges[e];
        // an edge only has to be looked at again if its source was lowered in the previous sweep or earlier in this one
        // (stamp = sweep of the last lowering; sweep 1 sees stamp 0 everywhere and relaxes every edge)
        if (STAMPED && stamp[ed.x] + 1u < sweep) continue;
        unsigned long long ls = label[ed.x];
        if (ls < label[ed.y]) { atomicMin(&label[ed.y], ls); stamp[ed.y] = sweep; any++; }
    }
    // number of lowerings of this sweep: the host switches to the frontier form when sweeps stop paying for themselves
    if (__any_sync(0xffffffffu, any != 0)) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) any += __shfl_xor_sync(0xffffffffu, any, o);
        if (lane_id() == 0) { sc->changed = 1; atomicAdd(&sc->n_lowered, any); }
    }
}

// ---- frontier form (hot loci: tens of sweeps over 10^7 edges, almost all of them wasted) ----
// The edge list is sorted by source once (radix sort of src << 32 | dst) and cut into rows; a round then relaxes only the
// out-edges of the UMIs whose label was lowered in the round before.  Same unique fixpoint as the sweeps (label[v] = the
// earliest-visited UMI that reaches v, directional.rs:30-54): every lowering of label[u] puts u on the next frontier, so
// no out-edge of u is left unrelaxed with u's final label.
// CSR by source without a sort: out-degree histogram, exclusive scan (scan.cuh), scatter.  The order of a row's
// neighbours is whatever the atomics produce — the fixpoint only depends on the edge SET.
__global__ void __launch_bounds__(256) csr_degree_kernel(const uint2 *__restrict__ edges, u64 n_edges, u32 *__restrict__ deg) {
    const u64 stride = (u64)gridDim.x * 256;
    for (u64 e = (u64)blockIdx.x * 256 + threadIdx.x; e < n_edges; e += stride) atomicAdd(&deg[edges[e].x], 1u);
}
struct DegreeOf { const u32 *deg; __device__ u32 operator()(u64 u) const { return deg[u]; } };
struct RowEmit {
    u32 *row_ptr, *fill; u64 n_rows;
    __device__ void operator()(u64 u, u32 v, u32 ex) const {
        row_ptr[u] = ex; fill[u] = ex;
        if (u == n_rows - 1) row_ptr[u + 1] = ex + v;
    }
};
__global__ void __launch_bounds__(256) csr_fill_kernel(const uint2 *__restrict__ edges, u64 n_edges, u32 *__restrict__ fill, u32 *__restrict__ col) {
    const u64 stride = (u64)gridDim.x * 256;
    for (u64 e = (u64)blockIdx.x * 256 + threadIdx.x; e < n_edges; e += stride) { const uint2 ed = edges[e]; col[atomicAdd(&fill[ed.x], 1u)] = ed.y; }
}
// first frontier after the plain sweeps: every UMI lowered in the last sweep
__global__ void __launch_bounds__(256) frontier_init_kernel(u32 n_unique, const u32 *__restrict__ stamp, u32 last_sweep,
                                                            u32 *__restrict__ fout, u32 *cnt_out) {
    u32 v = blockIdx.x * 256 + threadIdx.x;
    bool live = v < n_unique && stamp[v] >= last_sweep;
    u32 m = __ballot_sync(0xffffffffu, live);
    if (m) {
        u32 base = 0;
        if (lane_id() == 0) base = atomicAdd(cnt_out, (u32)__popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (live) fout[base + __popc(m & lanemask_lt())] = v;
    }
}
// one round: thread per frontier UMI (out-degree is small: at most 3*L neighbours per unit of k)
__global__ void __launch_bounds__(256) frontier_relax_kernel(const u32 *__restrict__ row_ptr, const u32 *__restrict__ col,
                                                             unsigned long long *label, u32 *stamp, const u32 *__restrict__ fin,
                                                             const u32 *__restrict__ cnt_in, u32 *__restrict__ fout, u32 *cnt_out, u32 round) {
    const u32 n = *cnt_in, stride = gridDim.x * 256;
    for (u32 i = blockIdx.x * 256 + threadIdx.x; i < n; i += stride) {
        const u32 u = fin[i];
        const unsigned long long lu = label[u];
        const u32 e1 = row_ptr[u + 1];
        for (u32 e = row_ptr[u]; e < e1; e++) {
            const u32 d = col[e];
            if (lu < label[d]) {
                const unsigned long long old = atomicMin(&label[d], lu);
                if (lu < old && atomicExch(&stamp[d], round) != round) {
                    // warp-aggregated append (lanes that reach this point together share one reservation)
                    const u32 act = __activemask();
                    const u32 leader = (u32)__ffs(act) - 1u;
                    u32 base = 0;
                    if (lane_id() == leader) base = atomicAdd(cnt_out, (u32)__popc(act));
                    base = __shfl_sync(act, base, leader);
                    fout[base + __popc(act & lanemask_lt())] = d;
                }
            }
        }
    }
}

// ---- two-phase fixpoint: mutual components first, then the contracted graph ----
// An edge (s,d) whose reverse (d,s) also passed the count rule is MUTUAL: s and d reach each other, so they end with the
// same label.  With -p 0.5 exactly the edges among frequency-1 UMIs are mutual (1 >= 2*1-1; f_s >= 2 f_d - 1 and
// f_d >= 2 f_s - 1 force f_s = f_d = 1), and in a hot locus those UMIs (sequencing errors, molecules seen once) form
// components tens of hops across — which is what makes plain propagation need tens of sweeps.
//   Phase A  connected components over the mutual edges with a lock-free union-find (one pass over the edge list: find with
//            path halving, the larger root is hooked under the smaller by atomicCAS; then one flatten pass).  Work does not
//            depend on the diameter.  The component's label = min over its members (atomicMin into the root's slot).
//   Phase B  the contracted graph is materialised once — every edge becomes (root[src], root[dst]), edges inside a
//            component vanish — and min-label propagation runs on it: one-way edges follow strictly falling frequency, so
//            its depth is ~log2(max frequency) and a handful of sweeps settle it.  Then every UMI takes its root's label.
// Same unique fixpoint (label[v] = earliest visited UMI that reaches v): all members of a mutual component reach each other,
// so they share the minimum over everything that reaches any of them.
__device__ __forceinline__ u32 uf_find(u32 *parent, u32 v) {
    u32 p = parent[v];
    while (p != v) {
        const u32 gp = parent[p];
        if (gp != p) parent[v] = gp;          // path halving: racy but benign (always an ancestor, ids fall along a path)
        v = p; p = gp;
    }
    return v;
}
__global__ void __launch_bounds__(256) uf_init_kernel(u32 n_unique, u32 *__restrict__ parent) {
    const u32 v = blockIdx.x * 256 + threadIdx.x;
    if (v < n_unique) parent[v] = v;
}
__global__ void __launch_bounds__(256) uf_union_kernel(const uint2 *__restrict__ edges, u64 n_edges, const i32 *__restrict__ freq,
                                                       const i32 *__restrict__ thr, u32 *parent) {
    const u64 stride = (u64)gridDim.x * 256;
    for (u64 e = (u64)blockIdx.x * 256 + threadIdx.x; e < n_edges; e += stride) {
        const uint2 ed = edges[e];
        if (ed.x > ed.y) continue;                            // a mutual pair is in the list in both directions: take one
        if (freq[ed.x] > thr[ed.y]) continue;                 // reverse edge did not pass the rule: one-way
        u32 ra = ed.x, rb = ed.y;
        for (;;) {
            ra = uf_find(parent, ra); rb = uf_find(parent, rb);
            if (ra == rb) break;
            if (ra < rb) { const u32 t = ra; ra = rb; rb = t; }
            if (atomicCAS(&parent[ra], ra, rb) == ra) break;  // hooked the larger root under the smaller
        }
    }
}
// flatten + component label: parent[v] = root, label[root] = min over the members
__global__ void __launch_bounds__(256) uf_flatten_kernel(u32 n_unique, u32 *parent, unsigned long long *label) {
    const u32 v = blockIdx.x * 256 + threadIdx.x;
    if (v >= n_unique) return;
    const u32 r = uf_find(parent, v);
    if (r != v) { parent[v] = r; atomicMin(&label[r], label[v]); }
}
// The contracted graph is materialised once: every edge becomes (comp[src], comp[dst]); edges inside a mutual component
// (most edges of a hot locus) disappear.  Phase B then touches only this list: no per-sweep comp gathers, no pass over
// all U labels.  *n_out is a device-resident count that the sweeps read (no host round trip).
__global__ void __launch_bounds__(256) contract_edges_kernel(const uint2 *__restrict__ edges, u64 n_edges, const u32 *__restrict__ comp,
                                                             uint2 *__restrict__ out, unsigned long long *n_out) {
    const u64 stride = (u64)gridDim.x * 256;
    const u64 rounds = (n_edges + stride - 1) / stride;          // every lane runs the same number of rounds (warp collectives)
    for (u64 it = 0; it < rounds; it++) {
        const u64 e = it * stride + (u64)blockIdx.x * 256 + threadIdx.x;
        bool live = false; u32 cs = 0, cd = 0;
        if (e < n_edges) { const uint2 ed = edges[e]; cs = comp[ed.x]; cd = comp[ed.y]; live = cs != cd; }
        const u32 m = __ballot_sync(0xffffffffu, live);
        if (m) {
            unsigned long long base = 0;
            if (lane_id() == 0) base = atomicAdd(n_out, (unsigned long long)__popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (live) out[base + __popc(m & lanemask_lt())] = make_uint2(cs, cd);
        }
    }
}
__global__ void __launch_bounds__(256) expand_labels_kernel(u32 n_unique, const u32 *__restrict__ comp, unsigned long long *label) {
    u32 v = blockIdx.x * 256 + threadIdx.x;
    if (v < n_unique) { u32 c = comp[v]; if (c != v) label[v] = label[c]; }
}

// keep[u] = (root(u) == u); root id = low 32 bits of the label
__global__ void __launch_bounds__(256) keep_from_label_kernel(u32 n_unique, const unsigned long long *__restrict__ label,
                                                              u8 *__restrict__ keep) {
    u32 u = blockIdx.x * 256 + threadIdx.x;
    if (u < n_unique) keep[u] = ((u32)label[u] == u) ? 1 : 0;
}

// ---- upstream adjacency (greedy MIS) ----
#define MIS_UNDECIDED 0
#define MIS_KEPT 1
#define MIS_REMOVED 2

__global__ void __launch_bounds__(256) mis_edge_kernel(const uint2 *__restrict__ edges, u64 n_edges,
                                                       const unsigned long long *__restrict__ prio, u8 *state, u8 *blocked) {
    u64 stride = (u64)gridDim.x * 256;
    for (u64 e = (u64)blockIdx.x * 256 + threadIdx.x; e < n_edges; e += stride) {
        uint2 ed = edges[e];
        if (prio[ed.x] < prio[ed.y] && state[ed.y] == MIS_UNDECIDED) {
            u8 s = state[ed.x];
            if (s == MIS_KEPT) state[ed.y] = MIS_REMOVED;
            else if (s == MIS_UNDECIDED) blocked[ed.y] = 1;
        }
    }
}
__global__ void __launch_bounds__(256) mis_node_kernel(u32 n_unique, u8 *state, u8 *blocked, DevScalars *sc) {
    u32 u = blockIdx.x * 256 + threadIdx.x;
    u32 any = 0;
    if (u < n_unique) {
        if (state[u] == MIS_UNDECIDED) {
            if (!blocked[u]) state[u] = MIS_KEPT;
            any = 1;                      // something was undecided at the start of this round
        }
        blocked[u] = 0;
    }
    if (__any_sync(0xffffffffu, any) && lane_id() == 0) sc->changed = 1;
}
// label of a removed UMI = earliest kept neighbour (the visit that removed it); kept UMIs label themselves
__global__ void __launch_bounds__(256) mis_label_kernel(const uint2 *__restrict__ edges, u64 n_edges,
                                                        const unsigned long long *__restrict__ prio, const u8 *__restrict__ state,
                                                        unsigned long long *label) {
    u64 stride = (u64)gridDim.x * 256;
    for (u64 e = (u64)blockIdx.x * 256 + threadIdx.x; e < n_edges; e += stride) {
        uint2 ed = edges[e];
        if (state[ed.x] == MIS_KEPT && state[ed.y] == MIS_REMOVED) atomicMin(&label[ed.y], prio[ed.x]);
    }
}
__global__ void __launch_bounds__(256) mis_keep_kernel(u32 n_unique, const u8 *__restrict__ state, u8 *__restrict__ keep) {
    u32 u = blockIdx.x * 256 + threadIdx.x;
    if (u < n_unique) keep[u] = state[u] == MIS_KEPT ? 1 : 0;
}

// ---- K7 emit_compact (src/deduplicate_sam.rs:227-231): kept representatives -> ascending read
// indices (canonical output order = input order).  A bitmap over reads makes the order free.
__global__ void __launch_bounds__(256) mark_kept_kernel(u32 n_unique, const u8 *__restrict__ keep, const u32 *__restrict__ rep_idx,
                                                        u32 *__restrict__ bitmap) {
    u32 u = blockIdx.x * 256 + threadIdx.x;
    if (u < n_unique && keep[u]) { u32 r = rep_idx[u]; atomicOr(&bitmap[r >> 5], 1u << (r & 31)); }
}
struct BitmapCount { const u32 *bm; __device__ u32 operator()(u64 w) const { return __popc(bm[w]); } };
// push-order position -> caller's read index (umigpu_push_reads first_read_index), chunk table on the device
struct ChunkMap {
    const u64 *start, *first; u32 n;
    const u32 *orig;      // optional: push-order position -> record number inside its chunk (BAM pushes drop filtered records)
    __device__ __forceinline__ u64 operator()(u64 r) const {
        u32 lo = 0, hi = n;
        while (hi - lo > 1) { u32 mid = (lo + hi) >> 1; if (start[mid] <= r) lo = mid; else hi = mid; }
        return first[lo] + (orig ? (u64)orig[r] : r - start[lo]);
    }
};
struct BitmapEmit {
    const u32 *bm; u64 *kept; u64 n_words; DevScalars *sc; ChunkMap cm;
    __device__ void operator()(u64 w, u32 cnt, u32 ex) const {
        u32 bits = bm[w];
        u32 o = ex;
        while (bits) { u32 b = __ffs(bits) - 1; bits &= bits - 1; kept[o++] = cm(w * 32 + b); }
        if (w == n_words - 1) sc->n_kept = ex + cnt;
    }
};

// per read: read index of the emitted representative of its cluster (ClusterTracker, --tag)
__global__ void __launch_bounds__(256) read_roots_kernel(u64 n, const u32 *__restrict__ read_uid,
                                                         const unsigned long long *__restrict__ label,
                                                         const u32 *__restrict__ rep_idx, ChunkMap cm, u64 *__restrict__ out,
                                                         u64 *__restrict__ out_umi_rep) {
    u64 i = (u64)blockIdx.x * 256 + threadIdx.x;
    if (i < n) {
        const u32 u = read_uid[i];
        out[i] = cm(rep_idx[(u32)label[u]]);          // representative of the read's cluster
        out_umi_rep[i] = cm(rep_idx[u]);              // representative of the read's own (bucket, UMI)
    }
}
