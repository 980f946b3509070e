// Key-record pass of the fused warp + mask + multi-agent group attention (the persistent tcgen05 kernel that consumes
// the records is csrc/attn_fa2.cuh).  Replaces
//   HeteroFusionBlock.warp_features            hetero_fusion.py:338-361   (geometry part)
//   get_roi_and_cav_mask / warp_affine          torch_transformation_utils.py:11-134, 254-355
//
//   tap_records_kernel   once per partition kind per FORWARD (the poses do not change between the block
//                        iterations): for every (scene b, ego i, token group g) the fp64 source-pixel map of every
//                        (source j, token) of the group, the bit-exact ROI visibility, and the COMPACTED list of
//                        visible keys as 16-byte records (row offset of the tap corner, 4 bilinear weights as bf16,
//                        source, slot, source type).  The ego's own keys are ordinary records (identity pose ->
//                        one tap of weight 1).
#pragma once
#include "attn_split.cuh"

namespace hmvit {

// MN-major operand (rows of 128 B = 64 consecutive MN elements for one k; 8-row groups 1024 B apart)
HMVIT_DEVINL uint64_t umma_desc_sw128_mn(uint32_t smem_addr) { return umma_desc_sw128(smem_addr); }

// one visible key of a (scene, ego, group): 16 bytes
struct KeyRec {
  uint32_t off;                // row offset of the footprint corner in the scene's K' / V' planes, in 16-byte units:
                               // ((source j * N + y0 * W + x0) * 32); the corner lies in the map with both neighbours
  uint32_t w01, w23;           // tap weights as bf16 pairs: (w00, w01), (w10, w11); 0 = tap that fell outside the map
  uint32_t meta;               // source j | slot << 8 | source type << 16 | relative-position offset << 24
};
static_assert(sizeof(KeyRec) == 16, "KeyRec must be 16 bytes");

constexpr int kRecMaxL = 8;    // agents per scene the record pass handles (2 x 256 entries per CTA)
constexpr int kFusedMaxAgents = 1024;   // B * L the persistent kernel's item table holds

struct RecParams {
  AttnParams a;                // geometry: B, L, H, W, mode, record_len, cav_mask, T, cell, key_mask
  KeyRec* rec;                 // [nkinds][B*L][G][L*64]
  int* nvis;                   // [nkinds][B*L][G] visible keys of the (ego, group)
  int kind0;                   // kind of grid.z == 0
};

// grid (G, B*L, nkinds)
__global__ void __launch_bounds__(256, 4) tap_records_kernel(const RecParams rp) {
  const AttnParams& p = rp.a;
  const int a = blockIdx.y;
  const int b = a / p.L, i = a - b * p.L;
  const int nrec = min(p.record_len[b], p.L);
  if (i >= nrec) return;
  const int kind = rp.kind0 + blockIdx.z;
  const int N = p.H * p.W;
  const int GX = p.W / kWin, G = (p.H / kWin) * GX;
  const int grp = blockIdx.x;
  const int gy = grp / GX, gx = grp - gy * GX;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  __shared__ int sCnt[kRecMaxL * 2];
  const int nent = nrec * kS;
  KeyRec rec[2];
  uint32_t bal[2];
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int e = it * 256 + threadIdx.x;
    KeyRec r; r.off = 0; r.w01 = 0; r.w23 = 0; r.meta = 0;
    bool vis = false;
    if (e < nent) {
      const int j = e >> 6, tk = e & 63;
      const int tj = p.mode[b * p.L + j] != 0 ? 1 : 0;
      r.meta = static_cast<uint32_t>(j) | (static_cast<uint32_t>(tk) << 8) | (static_cast<uint32_t>(tj) << 16) |
               (static_cast<uint32_t>((tk >> 3) * 15 + (tk & 7)) << 24);
      if (p.cav_mask[b * p.L + j] != 0) {
        const WarpMap wm = make_warp_map(p.T + ((static_cast<size_t>(b) * p.L + j) * p.L + i) * 16, p.H, p.W, p.cell);
        int rr, cc; group_token(kind, gy, gx, tk, p.H, p.W, rr, cc);
        double sx, sy; warp_src(wm, cc, rr, sx, sy);
        vis = warp_visible(sx, sy, p.H, p.W);
        if (p.key_mask != nullptr && p.key_mask[static_cast<size_t>(b * p.L + j) * N + rr * p.W + cc] == 0) vis = false;
        if (vis) {
          Taps tp = make_taps(sx, sy, p.H, p.W);
          // The footprint corner is moved INTO the map (a visible key has floor(sx) in [-1, W-1]): the taps that fell
          // outside carry weight 0 and now read in-map rows, so the gather can load all four taps unconditionally
          // (0 * finite == 0 exactly) -- no per-tap predicates, no zero-initialised destination registers.
          if (tp.x0 < 0) { tp.x0 = 0; tp.w00 = tp.w01; tp.w10 = tp.w11; tp.w01 = 0.f; tp.w11 = 0.f; }
          else if (tp.x0 > p.W - 2) { tp.x0 = p.W - 2; tp.w01 = tp.w00; tp.w11 = tp.w10; tp.w00 = 0.f; tp.w10 = 0.f; }
          if (tp.y0 < 0) { tp.y0 = 0; tp.w00 = tp.w10; tp.w01 = tp.w11; tp.w10 = 0.f; tp.w11 = 0.f; }
          else if (tp.y0 > p.H - 2) { tp.y0 = p.H - 2; tp.w10 = tp.w00; tp.w11 = tp.w01; tp.w00 = 0.f; tp.w01 = 0.f; }
          r.off = static_cast<uint32_t>((j * N + tp.y0 * p.W + tp.x0) * 32);
          r.w01 = pack_bf16x2(tp.w00, tp.w01); r.w23 = pack_bf16x2(tp.w10, tp.w11);
          // a visible key always has a non-zero tap weight (the nearest in-range corner has weight >= 1/4)
        }
      }
    }
    rec[it] = r;
    bal[it] = __ballot_sync(0xffffffffu, vis);
    if (lane == 0) sCnt[it * 8 + warp] = __popc(bal[it]);
  }
  __syncthreads();
  const size_t ag = (static_cast<size_t>(blockIdx.z) * p.B * p.L + a) * G + grp;
  KeyRec* dst = rp.rec + ag * (static_cast<size_t>(p.L) * kS);
  int nv = 0;
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int seg = it * 8 + warp;
    int base = 0;
#pragma unroll
    for (int s = 0; s < 2 * kRecMaxL; ++s) {
      const int c = sCnt[s];
      if (s < seg) base += c;
      nv += (it == 0) ? c : 0;
    }
    if ((bal[it] >> lane) & 1u)
      *reinterpret_cast<uint4*>(dst + base + __popc(bal[it] & ((1u << lane) - 1u))) = *reinterpret_cast<const uint4*>(&rec[it]);
  }
  if (threadIdx.x == 0) rp.nvis[ag] = nv;
}

// ------------------------------------------------------------------------------------------------
struct FusedAttnParams {
  AttnParams a;
  const KeyRec* rec;           // [B*L][G][L*64] records of a.kind
  const int* nvis;             // [B*L][G]
};

#ifndef HMVIT_FA_DBG         // bottleneck-hunting builds only (results are wrong): 1 no tap loads, 2 no softmax math, 4 no MMAs
#define HMVIT_FA_DBG 0
#endif

#ifdef HMVIT_TS   // timeline build (tools/fa_timeline.py): CTAs 0-1 record clock64() per role at pipeline events
__device__ unsigned long long g_fa_ts[2][4][1024];      // [cta][role: 0 softmax, 1 gather, 2 mma, 3 q-load][event]
#define FA_TS(role, idx, code) do { const int i_ = (idx); if (blockIdx.x < 2 && i_ < 1024) g_fa_ts[blockIdx.x][role][i_] = (static_cast<unsigned long long>(clock64()) << 8) | (code); } while (0)
#else
#define FA_TS(role, idx, code) do { } while (0)
#endif

// 16-byte read-only global load, issued in program order; pred == false: no load, zeros
HMVIT_DEVINL uint4 ldg_nc_u4_if(const uint4* p, bool pred) {
  uint4 v;
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t"
      "mov.b32 %0, 0;\n\tmov.b32 %1, 0;\n\tmov.b32 %2, 0;\n\tmov.b32 %3, 0;\n\t"
      "@p ld.global.nc.v4.b32 {%0, %1, %2, %3}, [%4];\n\t}"
      : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
      : "l"(p), "r"(pred ? 1 : 0));
  return v;
}
// unconditional 16-byte read-only global load, issued in program order
HMVIT_DEVINL uint4 ldg_nc_u4_v(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
HMVIT_DEVINL uint32_t hfma2_bf16_v(uint32_t a, uint32_t b, uint32_t c) {      // program-order variant of hfma2_bf16
  uint32_t d;
  asm volatile("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
HMVIT_DEVINL float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}

}  // namespace hmvit
