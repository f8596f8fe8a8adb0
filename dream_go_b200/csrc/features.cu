// features.cu -- V1 feature planes computed on the device from raw stones.
//
// Replaces the host's per-point work of `features::V1::get_features` (src/libdg_go/utils/features.rs:154-250) and the
// 361 x `Board::is_valid` of the prior construction (pool/policy_helper.rs:39-43) for the engine's self-play path: the
// host sends 384 bytes per leaf (stone / visited masks, hashes, last moves, and the two ladder planes, which stay a
// sequential search on the host) and gets the evaluation plus the legal-move mask back.
//
// One CTA per position, one thread per board point, everything in shared memory:
//   1. chains by min-label propagation with pointer jumping (board_fast.rs keeps linked lists; here they are rebuilt),
//   2. one 361-bit liberty set per chain (atomicOr from the empty points) and the XOR of its stones' zobrist keys,
//   3. per point the 32 plane bits: liberties of stones, legality + liberties-after-playing for both colours
//      (board_fast.rs:216-243, 484-539: union of the joined chains' liberty sets, the point's empty neighbours and any
//      captured stones that touch the new chain), super-ko against the last 16 hashes (board.rs:132-141),
//   4. written as `dg_packed_position` (index = symmetry[p]) plus the legal mask, and expanded in the same kernel into
//      the tower's input rows (400 x 64 fp16 per position).
// Integer work end to end; tests/test_features_gpu.py checks it bit for bit against the host code and the oracle.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"
#include "layout.h"

namespace dg {

struct RawPosition {                 // == dg_raw_position (include/dg_engine.h)
    uint32_t black[12], white[12], visited[12], ladder_capture[12], ladder_escape[12];
    unsigned long long hash;
    unsigned long long hash_history[16];
    int16_t last_move[2];
    uint16_t k_bits;
    uint8_t to_move, symmetry;
};
static_assert(sizeof(RawPosition) == 384, "dg_raw_position layout");

__constant__ unsigned long long c_zobrist[2][361];   // [colour - 1][point]
__constant__ uint16_t c_symmetry[8][361];

__device__ __forceinline__ bool bit(const uint32_t* m, int i) { return (m[i >> 5] >> (i & 31)) & 1u; }

// Block n < batch handles position n and also expands its planes into the tower's input rows (what pack_compact_kernel
// does for host-made planes: 400 rows x 64 fp16 channels, halo rows and channels 32..63 zero); block `batch` zeroes the
// rows between the last position and the end of the last 128-row tile (stale after a larger batch).
__global__ void __launch_bounds__(384) planes_from_stones_kernel(const RawPosition* __restrict__ in, uint32_t* __restrict__ out_planes,
                                                                 uint8_t* __restrict__ out_legal, uint4* __restrict__ out_rows,
                                                                 int batch, int total_rows) {
    __shared__ uint32_t planes_s[361];
    if (static_cast<int>(blockIdx.x) == batch) {
        const long first = static_cast<long>(batch) * DG_POS_ROWS * 8, last = static_cast<long>(total_rows) * 8;
        for (long i = first + threadIdx.x; i < last; i += 384) out_rows[static_cast<long>(DG_GUARD_ROWS) * 8 + i] = make_uint4(0, 0, 0, 0);
        return;
    }
    __shared__ uint32_t lib[361][12];             // liberty set of the chain whose smallest point index is the row
    __shared__ unsigned long long chash[361];     // XOR of the zobrist keys of its stones
    __shared__ uint16_t lab[384];
    __shared__ uint16_t nlib[361];
    __shared__ uint8_t col[384];
    __shared__ int changed, any_ko;

    const RawPosition& r = in[blockIdx.x];
    const int t = threadIdx.x;
    const bool on = t < 361;
    const int x = t % 19, y = t / 19;
    int nb[4];
    nb[0] = (on && x < 18) ? t + 1 : -1;
    nb[1] = (on && y > 0) ? t - 19 : -1;
    nb[2] = (on && x > 0) ? t - 1 : -1;
    nb[3] = (on && y < 18) ? t + 19 : -1;
    const int c = !on ? 0 : bit(r.black, t) ? 1 : bit(r.white, t) ? 2 : 0;
    col[t] = static_cast<uint8_t>(c);
    lab[t] = static_cast<uint16_t>(t);
    for (int i = t; i < 361 * 12; i += 384) (&lib[0][0])[i] = 0;
    if (on) chash[t] = 0;
    if (t == 0) any_ko = 0;
    __syncthreads();

    // 1. chains: every stone takes the smallest label among itself, its same-coloured neighbours and the label's label
    for (;;) {
        if (t == 0) changed = 0;
        __syncthreads();
        if (c) {
            int best = lab[t];
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (nb[k] >= 0 && col[nb[k]] == c) best = min(best, static_cast<int>(lab[nb[k]]));
            best = min(best, static_cast<int>(lab[best]));
            if (best < lab[t]) { lab[t] = static_cast<uint16_t>(best); changed = 1; }
        }
        __syncthreads();
        const int again = changed;
        __syncthreads();
        if (!again) break;
    }

    // 2. liberty sets and stone hashes per chain
    if (on) {
        if (c) atomicXor(&chash[lab[t]], c_zobrist[c - 1][t]);
        else {
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (nb[k] >= 0 && col[nb[k]]) atomicOr(&lib[lab[nb[k]]][t >> 5], 1u << (t & 31));
        }
    }
    __syncthreads();
    if (on && c && lab[t] == t) {
        int n = 0;
#pragma unroll
        for (int w = 0; w < 12; w++) n += __popc(lib[t][w]);
        nlib[t] = static_cast<uint16_t>(n);
    }
    __syncthreads();

    // 3. plane bits of this point
    const int tm = r.to_move, opp = 3 - tm;
    uint32_t m = 0;
    bool legal = false;
    if (on && c) {
        const int n = min(static_cast<int>(nlib[lab[t]]), 6);
        m = ((1u << n) - 1u) << (c == tm ? 5 : 17);
    } else if (on) {
        bool ko = false;
        int counts[2];
#pragma unroll
        for (int side_i = 0; side_i < 2; side_i++) {
            const int side = side_i == 0 ? tm : opp;
            uint32_t L[12];
#pragma unroll
            for (int w = 0; w < 12; w++) L[w] = 0;
            int n_empty = 0, nf = 0, nc = 0, fr[4], cap[4];
            bool ok = false;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int q = nb[k];
                if (q < 0) continue;
                const int cq = col[q];
                if (!cq) {
                    n_empty++;
#pragma unroll
                    for (int w = 0; w < 12; w++) if (w == (q >> 5)) L[w] |= 1u << (q & 31);
                    continue;
                }
                const int root = lab[q];
                const int n = nlib[root];
                if (cq == side) {
                    ok |= n >= 2;
                    bool dup = false;
                    for (int j = 0; j < nf; j++) dup |= fr[j] == root;
                    if (!dup) fr[nf++] = root;
                } else if (n == 1) {
                    bool dup = false;
                    for (int j = 0; j < nc; j++) dup |= cap[j] == root;
                    if (!dup) cap[nc++] = root;
                }
            }
            int count = -1;
            if (ok || n_empty || nc) {
                if (!nf && !nc) count = n_empty;
                else {
                    for (int j = 0; j < nf; j++)
#pragma unroll
                        for (int w = 0; w < 12; w++) L[w] |= lib[fr[j]][w];
#pragma unroll
                    for (int w = 0; w < 12; w++) if (w == (t >> 5)) L[w] &= ~(1u << (t & 31));
                    if (nc) {                    // rare: captured stones that touch the new chain become liberties
                        for (int s = 0; s < 361; s++) {
                            if (col[s] != 3 - side) continue;
                            const int root = lab[s];
                            bool is_cap = false;
                            for (int j = 0; j < nc; j++) is_cap |= cap[j] == root;
                            if (!is_cap) continue;
                            const int sx = s % 19, sy = s / 19;
                            const int sn[4] = {sx < 18 ? s + 1 : -1, sy > 0 ? s - 19 : -1, sx > 0 ? s - 1 : -1, sy < 18 ? s + 19 : -1};
                            bool touches = false;
                            for (int k = 0; k < 4; k++) {
                                const int u = sn[k];
                                if (u < 0) continue;
                                if (u == t) touches = true;
                                else if (col[u] == side)
                                    for (int j = 0; j < nf; j++) touches |= fr[j] == lab[u];
                            }
                            if (touches)
#pragma unroll
                                for (int w = 0; w < 12; w++) if (w == (s >> 5)) L[w] |= 1u << (s & 31);
                        }
                    }
                    count = 0;
#pragma unroll
                    for (int w = 0; w < 12; w++) count += __popc(L[w]);
                }
                if (side_i == 0 && bit(r.visited, t)) {  // super-ko: the position after the move is one of the last 16
                    unsigned long long h = r.hash ^ c_zobrist[tm - 1][t];
                    for (int j = 0; j < nc; j++) h ^= chash[cap[j]];
#pragma unroll
                    for (int j = 0; j < 16; j++) ko |= r.hash_history[j] == h;
                }
            }
            counts[side_i] = count;
        }
        if (counts[0] >= 0) m |= ((1u << min(counts[0], 6)) - 1u) << 11;
        if (counts[1] >= 0) m |= ((1u << min(counts[1], 6)) - 1u) << 23;
        if (ko) { m |= 1u << 29; any_ko = 1; }
        if (counts[0] >= 0) {
            if (bit(r.ladder_capture, t)) m |= 1u << 30;
            if (bit(r.ladder_escape, t)) m |= 1u << 31;
        }
        legal = counts[0] >= 0 && !ko;
    }
    if (on) {
        if (r.last_move[0] == t) m |= 1u << 3;
        if (r.last_move[1] == t) m |= 1u << 4;
    }
    __syncthreads();
    // 4. output
    if (on) {
        const uint32_t global = (tm == 1 ? 1u : 2u) | (any_ko ? 4u : 0u);
        uint32_t* out = out_planes + static_cast<size_t>(blockIdx.x) * 362;
        const int target = c_symmetry[r.symmetry][t];
        out[target] = m | global;
        planes_s[target] = m | global;
        if (t == 0) out[361] = r.k_bits;
        out_legal[static_cast<size_t>(blockIdx.x) * 361 + t] = legal ? 1 : 0;
    }
    __syncthreads();
    // 5. the tower's input rows of this position: one 16-byte chunk (8 channels) per thread and step, coalesced
    const uint32_t kbits = r.k_bits;
    uint4* rows = out_rows + (static_cast<long>(DG_GUARD_ROWS) + static_cast<long>(blockIdx.x) * DG_POS_ROWS) * 8;
    for (int idx = t; idx < DG_POS_ROWS * 8; idx += 384) {
        const int q = idx >> 3, chunk = idx & 7;
        const int qx = q % DG_LINE_STRIDE, qy = q / DG_LINE_STRIDE;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (chunk < 4 && qx < 19 && qy < 19) {
            const uint32_t mask = planes_s[qy * 19 + qx] >> (chunk * 8);
            uint32_t h[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const uint32_t one = (chunk == 0 && j < 2) ? kbits : 0x3c00u;
                h[j] = ((mask >> j) & 1u) ? one : 0u;
            }
            v = make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), h[6] | (h[7] << 16));
        }
        rows[idx] = v;
    }
}

cudaError_t upload_feature_tables(const unsigned long long* zobrist /* [2][361] */, const uint16_t* symmetry /* [8][361] */) {
    cudaError_t rc = cudaMemcpyToSymbol(c_zobrist, zobrist, sizeof(c_zobrist));
    if (rc != cudaSuccess) return rc;
    return cudaMemcpyToSymbol(c_symmetry, symmetry, sizeof(c_symmetry));
}

cudaError_t launch_planes_from_stones(const void* raw, void* planes, void* legal, void* rows64, int batch, cudaStream_t s) {
    const int total = dg_num_tiles(batch) * DG_TILE_M;
    planes_from_stones_kernel<<<batch + 1, 384, 0, s>>>(static_cast<const RawPosition*>(raw), static_cast<uint32_t*>(planes),
                                                        static_cast<uint8_t*>(legal), static_cast<uint4*>(rows64), batch, total);
    return cudaGetLastError();
}

}  // namespace dg
