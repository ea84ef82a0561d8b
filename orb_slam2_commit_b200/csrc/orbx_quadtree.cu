// orbx_quadtree.cu — level-synchronous, bit-exact parallel form of ORBextractor::DistributeOctTree
// (ORBextractor.cc:562-815) and ExtractorNode::DivideNode (:501-560).
//
// One CTA per (level, frame): 256 threads (several CTAs per SM) for frames up to 1 Mpx, 1024 above. The reference keeps a std::list of nodes, each owning a vector of
// keypoints, and splits nodes one at a time; here the list is an ARRAY IN LIST ORDER held in shared memory
// (bounds + count per node) and every candidate keypoint carries the index of the node that currently owns it
// (u16 in HBM/L2). One "pass" = what the reference does in one sweep over its list:
//   full pass     every node with more than one keypoint is split (reference :628-712)
//   careful pass  the splittable nodes are sorted by (count desc, creation order desc) and split one by one until
//                 the list reaches N (reference :720-788). The reference breaks count ties by comparing list-node
//                 ADDRESSES; this build (and the oracle) use creation order, see DESIGN.md.
// Children are pushed to the FRONT of the reference's list in n1..n4 order while parents are visited front to back,
// so the new list is [children in reverse creation order] ++ [untouched nodes in old order]; new positions come from
// two block-wide exclusive scans. A keypoint sweep then moves every keypoint to its child (two integer compares
// against the parent's midpoint, exactly DivideNode's float compares because all coordinates are small integers)
// and histograms the NEXT split (warp-aggregated shared-memory atomics), so each pass reads the candidates once.
// Finally every surviving node emits its best keypoint: max response, first in list order on ties (:796-812).
#include "orbx_internal.cuh"
#include <algorithm>
#include <cstdlib>

#define QT_MAX ORBX_QT_THREADS   // the kernel runs with blockDim.x = 256 (small frames) or 1024 (large frames)

struct QtBounds { short x0, x1, y0, y1; };

__device__ __forceinline__ int qt_quadrant(const QtBounds b, const int x, const int y)
{
    const int midx = b.x0 + ((b.x1 - b.x0 + 1) >> 1);   // UL.x + ceil((UR.x-UL.x)/2)
    const int midy = b.y0 + ((b.y1 - b.y0 + 1) >> 1);
    return (x >= midx ? 1 : 0) + (y >= midy ? 2 : 0);    // n1,n2,n3,n4 -> 0,1,2,3
}

// exclusive scan of vals[0..n) in place; returns the total. Every thread of the CTA must call it.
__device__ int qt_scan(int* vals, const int n, int* s_w)
{
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int QT = blockDim.x, nwarps = QT >> 5;
    int carry = 0;
    for (int base = 0; base < n; base += QT) {
        const int i = base + tid;
        const int v = i < n ? vals[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) s_w[wid] = x;
        __syncthreads();
        if (wid == 0) {
            const int t = lane < nwarps ? s_w[lane] : 0;
            int z = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, z, o); if (lane >= o) z += y; }
            s_w[lane] = z - t;
            if (lane == 31) s_w[32] = z;
        }
        __syncthreads();
        if (i < n) vals[i] = carry + s_w[wid] + x - v;
        carry += s_w[32];
        __syncthreads();
    }
    return carry;
}

// warp-aggregated atomicAdd(+1) on shared counters: lanes hitting the same counter elect one leader
__device__ __forceinline__ void qt_count(int* counters, const int key, const bool active)
{
    const unsigned peers = __match_any_sync(0xffffffffu, active ? key : -1);
    if (active && (__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&counters[key], __popc(peers));
}

// SMALL: 256-thread CTAs, several per SM (they hide each other's barriers); otherwise one 1024-thread CTA per SM
template <bool SMALL>
__global__ void __launch_bounds__(SMALL ? 256 : QT_MAX, SMALL ? 8 : 1) quadtree_kernel(OrbxFrameLayout L)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int tid = threadIdx.x;
    const int QT = blockDim.x, nwarps = QT >> 5;
    const int level = blockIdx.x, frame = blockIdx.y;
    const OrbxLevelGeom g = L.lvl[level];
    const int C = L.qt_cap, N = g.quota;

    // ---- shared-memory carve-up
    QtBounds* nb[2];
    int* ncnt[2];
    int* child[2];
    nb[0] = reinterpret_cast<QtBounds*>(smem_raw);
    nb[1] = nb[0] + C;
    ncnt[0] = reinterpret_cast<int*>(nb[1] + C);
    ncnt[1] = ncnt[0] + C;
    child[0] = ncnt[1] + C;
    child[1] = child[0] + 4 * C;
    int* vpre = child[1] + 4 * C;      // per node: #non-empty children -> exclusive prefix in visiting order
    int* keep = vpre + C;              // per node: survives untouched -> new index
    unsigned char* divf = reinterpret_cast<unsigned char*>(keep + C);   // per node: split in this pass
    int sortn = 1; while (sortn < C) sortn <<= 1;
    unsigned long long* skey = reinterpret_cast<unsigned long long*>(
        (reinterpret_cast<uintptr_t>(divf + C) + 7) & ~(uintptr_t)7);   // [sortn]
    int* sc = reinterpret_cast<int*>(skey + sortn);                      // [sortn] scratch for the careful pass
    int* ndc[2];                                                         // fast path: node = depth << 24 | code
    ndc[0] = sc + sortn;
    ndc[1] = ndc[0] + C;
    int* H = ndc[1] + C;                                                 // fast path: per-depth count pyramid
    __shared__ int s_w[33];
    __shared__ int s_misc[4];

    uint32_t* cand = L.cand + (size_t)frame * L.cand_total + g.cand_off;
    uint16_t* node = L.cand_node + (size_t)frame * L.cand_total + g.cand_off;

    // ---- 0. gather the cells' survivors into the reference's list order (cell-row-major, FAST row-major inside)
    int n = 0;
    {
        const int ncl = g.ncols * g.nrows;
        const int* cc = L.cell_count + (size_t)frame * L.ncells + g.cell0;
        const uint32_t* slots = L.slots + (size_t)frame * L.slot_total;
        for (int base = 0; base < ncl; base += QT) {
            const int ci = base + tid;
            const int cnt = ci < ncl ? cc[ci] : 0;
            // block exclusive scan of cnt (one chunk)
            const int lane = tid & 31, wid = tid >> 5;
            int x = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
            if (lane == 31) s_w[wid] = x;
            __syncthreads();
            if (wid == 0) {
                const int t = lane < nwarps ? s_w[lane] : 0;
                int z = t;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, z, o); if (lane >= o) z += y; }
                s_w[lane] = z - t;
                if (lane == 31) s_w[32] = z;
            }
            __syncthreads();
            // the copy itself is done warp-per-cell (coalesced reads of a cell's slot list, coalesced writes, several
            // cells in flight per warp) instead of thread-per-cell: offsets / counts / slot addresses of this chunk of
            // cells go through shared memory (the node arrays are not in use yet)
            int* g_off = reinterpret_cast<int*>(smem_raw);
            int* g_cnt = g_off + QT;
            int* g_slot = g_cnt + QT;
            g_off[tid] = n + s_w[wid] + x - cnt;
            g_cnt[tid] = cnt;
            g_slot[tid] = cnt > 0 ? L.cells[g.cell0 + ci].slot_off : 0;
            __syncthreads();
            const int nc = min(QT, ncl - base);
#pragma unroll 4
            for (int c = wid; c < nc; c += nwarps) {
                const int k = g_cnt[c], o = g_off[c];
                const uint32_t* sl = slots + g_slot[c];
                for (int j = lane; j < k; j += 32) if (o + j < g.cand_cap) cand[o + j] = __ldg(sl + j);
            }
            n += s_w[32];
            __syncthreads();
        }
        if (n > g.cand_cap) n = g.cand_cap;
        if (tid == 0) L.cand_count[(size_t)frame * L.nlevels + level] = n;
    }
    __syncthreads();

    // ---- 1+2. Two ways to run the reference's passes.
    // FAST PATH (no keypoint sweep per pass): DivideNode's geometry is a pure function of (root, quadrant path), so one
    // sweep computes every keypoint's path code down to depth DEPTH and histograms the deepest level; the counts of
    // all shallower nodes follow by summing children (a count pyramid in shared memory). The passes then run on the
    // node list alone: a node is (depth, code), its children's counts are pyramid look-ups. If a node at the deepest
    // tracked depth ever has to be split (strongly clustered keypoints) the tree falls back to the
    // SWEEP PATH: keypoints carry their node index and every pass moves them to their child and histograms the next
    // split (one sweep per pass).
    const int DEPTH = g.qt_depth;
    bool fast = DEPTH > 0;
    int cur = 0, S = 0;
    auto hbase = [&](int d) { return g.nini * (((1 << (2 * d)) - 1) / 3); };   // first entry of depth d in H
    for (;;) {
    int mode;                   // -1 root filter (sweep path only), 0 full pass, 1 careful pass
    cur = 0;
    if (fast) {
        const int T = hbase(DEPTH + 1);
        for (int i = tid; i < T; i += QT) H[i] = 0;
        __syncthreads();
        int* HD = H + hbase(DEPTH);
        for (int i0 = 0; i0 < n; i0 += QT) {
            const int i = i0 + tid;
            int code = -1;
            if (i < n) {
                // DivideNode's quadrant path is separable: the x choices depend on x alone (root included), the y choices on
                // y alone, so the host lays out both halves of the code per coordinate (orbx_qt_path_tables): two loads and
                // an OR instead of DEPTH rounds of midpoint arithmetic (29 % of this kernel's instructions before)
                const uint32_t xy = cand[i];
                code = (int)(__ldg(L.qt_path + g.qt_xs_off + (xy & 0xfff)) | __ldg(L.qt_path + g.qt_ys_off + ((xy >> 12) & 0xfff)));
                node[i] = (uint16_t)code;
            }
            // plain shared-memory atomics: consecutive candidates often share a deepest-level node, but the native atomic
            // serialises those few lanes faster than __match_any_sync finds them (0.279 -> 0.271 ms per 512 frames)
            if (i < n) atomicAdd(&HD[code], 1);
        }
        __syncthreads();
        for (int d = DEPTH - 1; d >= 0; d--) {
            int* Hd = H + hbase(d);
            const int* Hc = H + hbase(d + 1);
            for (int c = tid; c < (g.nini << (2 * d)); c += QT) Hd[c] = Hc[4 * c] + Hc[4 * c + 1] + Hc[4 * c + 2] + Hc[4 * c + 3];
            __syncthreads();
        }
        // list = non-empty roots in root order
        for (int r = tid; r < g.nini; r += QT) keep[r] = H[r] > 0;
        __syncthreads();
        S = qt_scan(keep, g.nini, s_w);
        for (int r = tid; r < g.nini; r += QT)
            if (H[r] > 0) { ndc[0][keep[r]] = r; ncnt[0][keep[r]] = H[r]; }
        __syncthreads();
        mode = 0;
    } else {
        // roots (ORBextractor.cc:567-626): nIni nodes of width hX; keypoint -> root (int)(x / hX)
        S = g.nini;
        for (int r = tid; r < S; r += QT) {
            QtBounds b;
            b.x0 = (short)(int)(g.hx * (float)r);
            b.x1 = (short)(int)(g.hx * (float)(r + 1));
            b.y0 = 0;
            b.y1 = (short)(g.h - 2 * ORBX_MINB);
            nb[cur][r] = b;
            ncnt[cur][r] = 0;
        }
        __syncthreads();
        for (int i0 = 0; i0 < n; i0 += QT) {
            const int i = i0 + tid;
            int r = -1;
            if (i < n) {
                r = (int)__fdiv_rn((float)(cand[i] & 0xfffu), g.hx);
                r = r < S ? r : S - 1;
                node[i] = (uint16_t)r;
            }
            qt_count(ncnt[cur], r, i < n);
        }
        __syncthreads();
        mode = -1;              // pass -1 only drops empty roots; then the reference's while(!bFinish) loop
    }

    bool finish = false, overflow = false;
    while (!finish) {
        if (fast) {
            // children's counts of every splittable node from the pyramid; a splittable node at the deepest tracked
            // depth means the pyramid is too shallow for this tree
            if (tid == 0) s_misc[3] = 0;
            __syncthreads();
            for (int i = tid; i < S; i += QT) {
                if (ncnt[cur][i] > 1) {
                    const int d = ndc[cur][i] >> 24, c = ndc[cur][i] & 0xffffff;
                    if (d >= DEPTH) s_misc[3] = 1;
                    else {
                        const int* Hc = H + hbase(d + 1) + 4 * c;
                        int* ch = child[cur] + 4 * i;
                        ch[0] = Hc[0]; ch[1] = Hc[1]; ch[2] = Hc[2]; ch[3] = Hc[3];
                    }
                }
            }
            __syncthreads();
            if (s_misc[3]) { overflow = true; break; }
        }
        const int nxt = cur ^ 1;
        int Ctot = 0;           // children created in this pass
        // (A) which nodes are split, and in which order they are visited
        if (mode <= 0) {
            for (int i = tid; i < S; i += QT) {
                const int c = ncnt[cur][i];
                int nz = 0;
                if (mode == 0 && c > 1) {
                    const int* ch = child[cur] + 4 * i;
                    nz = (ch[0] > 0) + (ch[1] > 0) + (ch[2] > 0) + (ch[3] > 0);
                }
                divf[i] = nz > 0;
                vpre[i] = nz;
            }
            __syncthreads();
            Ctot = qt_scan(vpre, S, s_w);       // parents are visited in list order (front to back)
        } else {
            // careful pass: E candidates sorted by (count desc, creation desc == list index asc)
            for (int i = tid; i < S; i += QT) { keep[i] = ncnt[cur][i] > 1; divf[i] = 0; vpre[i] = 0; }
            __syncthreads();
            const int E = qt_scan(keep, S, s_w);
            int P = 1; while (P < E) P <<= 1;
            for (int j = tid; j < P; j += QT) skey[j] = 0ull;
            __syncthreads();
            for (int i = tid; i < S; i += QT)
                if (ncnt[cur][i] > 1)
                    skey[keep[i]] = ((unsigned long long)(unsigned)ncnt[cur][i] << 32) | (unsigned)(0xffffffffu - (unsigned)i);
            __syncthreads();
            for (int k = 2; k <= P; k <<= 1)                 // bitonic sort, descending
                for (int j = k >> 1; j > 0; j >>= 1) {
                    for (int t = tid; t < P; t += QT) {
                        const int u = t ^ j;
                        if (u > t) {
                            const unsigned long long a = skey[t], b = skey[u];
                            const bool desc = (t & k) == 0;
                            if (desc ? a < b : a > b) { skey[t] = b; skey[u] = a; }
                        }
                    }
                    __syncthreads();
                }
            // gains in sorted order, running size, first position where the list reaches N
            if (tid == 0) s_misc[0] = E;
            for (int j = tid; j < E; j += QT) {
                const int i = (int)(0xffffffffu - (unsigned)(skey[j] & 0xffffffffull));
                const int* ch = child[cur] + 4 * i;
                sc[j] = (ch[0] > 0) + (ch[1] > 0) + (ch[2] > 0) + (ch[3] > 0);   // children of candidate j
            }
            __syncthreads();
            const int allc = qt_scan(sc, E, s_w);            // sc[j] = children created before candidate j
            (void)allc;
            for (int j = tid; j < E; j += QT) {
                const int i = (int)(0xffffffffu - (unsigned)(skey[j] & 0xffffffffull));
                const int* ch = child[cur] + 4 * i;
                const int cj = (ch[0] > 0) + (ch[1] > 0) + (ch[2] > 0) + (ch[3] > 0);
                // list size after splitting candidates 0..j : S + sum(c - 1)
                if (S + sc[j] + cj - (j + 1) >= N) atomicMin(&s_misc[0], j);
            }
            __syncthreads();
            const int jstop = s_misc[0];
            const int D = jstop < E ? jstop + 1 : E;         // candidates actually split
            if (tid == 0) s_misc[1] = 0;
            __syncthreads();
            for (int j = tid; j < D; j += QT) {
                const int i = (int)(0xffffffffu - (unsigned)(skey[j] & 0xffffffffull));
                divf[i] = 1;
                vpre[i] = sc[j];
                if (j == D - 1) {
                    const int* ch = child[cur] + 4 * i;
                    s_misc[1] = sc[j] + (ch[0] > 0) + (ch[1] > 0) + (ch[2] > 0) + (ch[3] > 0);
                }
            }
            __syncthreads();
            Ctot = s_misc[1];
        }
        // (B) new list = children (reverse creation order) ++ untouched nodes (old order)
        for (int i = tid; i < S; i += QT) keep[i] = (!divf[i] && ncnt[cur][i] > 0) ? 1 : 0;
        __syncthreads();
        const int nkeep = qt_scan(keep, S, s_w);
        const int Snew = Ctot + nkeep;
        if (tid == 0) s_misc[2] = 0;
        __syncthreads();
        int myexp = 0;
        for (int i = tid; i < S; i += QT) {
            if (divf[i]) {
                const QtBounds b = nb[cur][i];
                const int midx = b.x0 + ((b.x1 - b.x0 + 1) >> 1), midy = b.y0 + ((b.y1 - b.y0 + 1) >> 1);
                int* ch = child[cur] + 4 * i;
                int r = 0;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int c = ch[q];
                    if (c > 0) {
                        const int ni = Ctot - 1 - (vpre[i] + r);
                        r++;
                        QtBounds cb;
                        cb.x0 = (q & 1) ? (short)midx : b.x0;  cb.x1 = (q & 1) ? b.x1 : (short)midx;
                        cb.y0 = (q & 2) ? (short)midy : b.y0;  cb.y1 = (q & 2) ? b.y1 : (short)midy;
                        if (ni < C) {
                            nb[nxt][ni] = cb; ncnt[nxt][ni] = c;
                            if (fast) ndc[nxt][ni] = (((ndc[cur][i] >> 24) + 1) << 24) | ((ndc[cur][i] & 0xffffff) * 4 + q);
                        }
                        myexp += c > 1;
                        ch[q] = ni;                          // remap table for the keypoint sweep
                    }
                }
            } else if (ncnt[cur][i] > 0) {
                const int ni = Ctot + keep[i];
                if (ni < C) { nb[nxt][ni] = nb[cur][i]; ncnt[nxt][ni] = ncnt[cur][i]; if (fast) ndc[nxt][ni] = ndc[cur][i]; }
                keep[i] = ni;
            }
        }
        if (myexp) atomicAdd(&s_misc[2], myexp);
        if (!fast) for (int i = tid; i < 4 * C; i += QT) child[nxt][i] = 0;
        __syncthreads();
        const int Enew = s_misc[2];
        const int Sn = Snew < C ? Snew : C;
        // (C) keypoint sweep: move to the child, histogram the next split (sweep path only)
        for (int i0 = 0; i0 < (fast ? 0 : n); i0 += QT) {
            const int i = i0 + tid;
            int key = -1;
            bool act = false;
            if (i < n) {
                const int old = node[i];
                const uint32_t xy = cand[i];
                const int x = xy & 0xfff, y = (xy >> 12) & 0xfff;
                int nw;
                if (divf[old]) nw = child[cur][4 * old + qt_quadrant(nb[cur][old], x, y)];
                else nw = keep[old];
                nw = nw < C ? nw : C - 1;
                node[i] = (uint16_t)nw;
                if (ncnt[nxt][nw] > 1) { key = 4 * nw + qt_quadrant(nb[nxt][nw], x, y); act = true; }
            }
            qt_count(child[nxt], key, act);
        }
        __syncthreads();
        // (D) the reference's loop control (:714-789)
        const int prevS = S;
        S = Sn;
        cur = nxt;
        if (mode == -1) mode = 0;
        else if (mode == 0) {
            if (S >= N || S == prevS) finish = true;
            else if (S + 3 * Enew > N) mode = 1;
        } else {
            if (S >= N || S == prevS) finish = true;
        }
    }

    if (overflow) { fast = false; continue; }     // pyramid too shallow: redo this tree on the sweep path
    break;
    }

    // ---- 3. best keypoint per node: max response, first in the reference's vector order on ties (:796-812)
    int* best = child[cur ^ 1];
    for (int i = tid; i < S; i += QT) best[i] = 0;
    if (fast)                                      // leaf (depth, code) -> -(list index + 1) in the count pyramid
        for (int i = tid; i < S; i += QT) H[hbase(ndc[cur][i] >> 24) + (ndc[cur][i] & 0xffffff)] = -(i + 1);
    __syncthreads();
    for (int i = tid; i < n; i += QT) {
        const uint32_t xy = cand[i];
        int idx = node[i];
        if (fast) {                                // walk the keypoint's path UP to the leaf that owns it: the pyramid is one level
            int code = idx, hb = hbase(DEPTH);     // deeper than a uniform spread needs, so the leaf is a step or two away
            idx = 0;                               // (entries below a leaf hold counts >= 0, the leaf itself -(index + 1))
            for (int t = DEPTH; t >= 0; t--) {
                const int v = H[hb + code];
                if (v < 0) { idx = -v - 1; break; }
                code >>= 2; hb = (hb - g.nini) >> 2;       // hbase(t - 1)
            }
        }
        atomicMax(&best[idx], (int)(((xy >> 24) << 23) | (0x7fffffu - (uint32_t)i)));
    }
    __syncthreads();
    uint32_t* out = L.lvl_kp + (size_t)frame * L.kp_cap_total + L.lvl_kp_off[level];
    for (int i = tid; i < S && i < g.kp_cap; i += QT) {
        const uint32_t xy = cand[0x7fffffu - ((uint32_t)best[i] & 0x7fffffu)];
        out[i] = (xy & 0xff000000u) | ((((xy >> 12) & 0xfffu) + ORBX_MINB) << 12) | ((xy & 0xfffu) + ORBX_MINB);
    }
    if (tid == 0) L.lvl_kp_count[(size_t)frame * L.nlevels + level] = S < g.kp_cap ? S : g.kp_cap;
}

// Host: per-coordinate halves of the depth-`g.qt_depth` path code of DivideNode (ORBextractor.cc:501-560), in the kernel's own
// integer arithmetic: code = root * 4^D + sum_d (qx_d + 2 qy_d) * 4^(D-1-d); xs[x] carries the root and the qx bits, ys[y] the
// qy bits. x, y are relative to minBorder (the candidates' coordinates).
void orbx_qt_path_tables(OrbxLevelGeom& g, std::vector<uint16_t>& out)
{
    const int D = g.qt_depth, wlen = g.w - 2 * ORBX_MINB, hlen = g.h - 2 * ORBX_MINB;
    g.qt_xs_off = (int)out.size();
    for (int x = 0; x <= std::max(wlen, 0); x++) {
        int r = (int)((float)x / g.hx);
        r = r < g.nini ? r : g.nini - 1;
        int x0 = (int)(g.hx * (float)r), x1 = (int)(g.hx * (float)(r + 1)), code = r;
        for (int d = 0; d < D; d++) {
            const int midx = x0 + ((x1 - x0 + 1) >> 1);
            const int qx = x >= midx;
            if (qx) x0 = midx; else x1 = midx;
            code = code * 4 + qx;
        }
        out.push_back((uint16_t)code);
    }
    g.qt_ys_off = (int)out.size();
    for (int y = 0; y <= std::max(hlen, 0); y++) {
        int y0 = 0, y1 = hlen, code = 0;
        for (int d = 0; d < D; d++) {
            const int midy = y0 + ((y1 - y0 + 1) >> 1);
            const int qy = y >= midy;
            if (qy) y0 = midy; else y1 = midy;
            code = code * 4 + 2 * qy;
        }
        out.push_back((uint16_t)code);
    }
}

static size_t qt_smem_bytes(int C, int hist_ints)
{
    int sortn = 1; while (sortn < C) sortn <<= 1;
    size_t b = 0;
    b += 2 * (size_t)C * sizeof(QtBounds);
    b += 2 * (size_t)C * sizeof(int);
    b += 2 * 4 * (size_t)C * sizeof(int);
    b += 2 * (size_t)C * sizeof(int);
    b += (size_t)C + 8;
    b += (size_t)sortn * (sizeof(unsigned long long) + sizeof(int));
    b += 2 * (size_t)C * sizeof(int) + (size_t)hist_ints * sizeof(int);
    b = std::max(b, (size_t)3 * QT_MAX * sizeof(int));     // the candidate gather borrows 3 ints per thread up front
    return (b + 15) & ~(size_t)15;
}

void orbx_launch_quadtree(const OrbxFrameLayout& L, int nframes, int threads, cudaStream_t st)
{
    const size_t smem = qt_smem_bytes(L.qt_cap, L.qt_hist_ints);
    static OrbxSmemMark mark[2] = {};
    orbx_need_smem(quadtree_kernel<true>, mark[0], smem);
    orbx_need_smem(quadtree_kernel<false>, mark[1], smem);
    dim3 grid(L.nlevels, nframes);
    if (threads == 256) quadtree_kernel<true><<<grid, 256, smem, st>>>(L);
    else quadtree_kernel<false><<<grid, QT_MAX, smem, st>>>(L);
}
