// Sampled-candidate evaluation on the device: the reference's evaluation()
// (utils.py:544-602) ranks one held-out item against 100 uniform negatives per user with a batch-1 predict() call
// and two argsorts per user.  Here:
//   sample_candidates : cand[u, 0] = held-out item, cand[u, 1..C) = uniform ids in 1..itemnum outside the user's
//                       train set (the rejection loop of utils.py:578-583), one thread per (user, slot)
//   candidate_rank    : logits[u, c] = <feats[u, :D], E[cand[u, c]]> (+ the SRFRN user term, SRFR_model.py:244-257),
//                       rank[u] = #{c >= 1 : logits[u, c] > logits[u, 0]}  == (-logits).argsort().argsort()[0] with
//                       ties resolved in favour of the held-out item (stable order, utils.py:591)
//   add_user_term     : logits[u, :] += <feats[u, D:D+F], Fe[label[u]]>   (SRFRN.predict's fake-label columns)
// HBM-bound gathers: one warp per user, a candidate row is read with 16-byte vectors by consecutive lanes.
#include "common.cuh"
#include "srfrd_b200.h"

namespace srfrd {

struct CandParams {
  const int64_t* offsets;   // (usernum + 1) CSR of the TRAIN interactions
  const int* items;         // (nnz)
  const int* users0;        // (U) 0-based user rows
  const int* target;        // (usernum) held-out item per user row
  int itemnum, C;
  int64_t U;
  uint64_t seed;
  int64_t* cand;            // (U, C)
};

__global__ void __launch_bounds__(256) sample_candidates_kernel(CandParams p) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.U * p.C) return;
  const int64_t u = i / p.C;
  const int c = (int)(i - u * p.C);
  const int u0 = p.users0[u];
  if (c == 0) { p.cand[i] = p.target[u0]; return; }
  const int64_t a = p.offsets[u0];
  const int n = (int)(p.offsets[u0 + 1] - a);
  int pick = 1;
  for (uint32_t tries = 0;; ++tries) {           // `while t in rated` (utils.py:581): loops until a free id is found
    pick = 1 + (int)(hash_u32(p.seed, 0xCA4Du + tries, (uint64_t)i) % (uint32_t)p.itemnum);
    bool hit = false;
    for (int k = 0; k < n; ++k) hit |= p.items[a + k] == pick;
    if (!hit || tries >= 4096) break;            // (a user who rated the whole catalogue cannot be served; bounded)
  }
  p.cand[i] = pick;
}

struct RankParams {
  const float* feats; int ldf;       // (U, ldf) fp32 sequence representations
  const float* table; int D;         // (n_rows, D) fp32 item table
  int64_t n_rows;
  const int64_t* cand;               // (U, C)
  const float* fake_table; int F;    // SRFRN: (3, F) or null
  const int64_t* user_label;         // SRFRN: (U) or null
  int64_t U; int C;
  float* logits; int ldl;            // (U, ldl) or null
  int* rank;                         // (U) or null
  int* err;                          // set to 1 when a candidate id is out of range (clamped to row 0)
};

__global__ void __launch_bounds__(256) candidate_rank_kernel(RankParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t u = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (u >= p.U) return;
  const float* f = p.feats + u * p.ldf;
  float user_term = 0.f;
  if (p.fake_table) {
    const int64_t lab = p.user_label[u];
    float s = 0.f;
    for (int c = lane; c < p.F; c += 32) s = fmaf(f[p.D + c], __ldg(p.fake_table + lab * p.F + c), s);
    user_term = warp_sum(s);
  }
  const bool vec = (p.D % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.table) & 15) == 0) && (p.ldf % 4 == 0) &&
                   ((reinterpret_cast<uintptr_t>(p.feats) & 15) == 0);
  float l0 = 0.f;
  int rank = 0;
  for (int c = 0; c < p.C; ++c) {
    int64_t id = __ldg(p.cand + u * p.C + c);
    if (id < 0 || id >= p.n_rows) { if (lane == 0 && p.err) *p.err = 1; id = 0; }
    const float* e = p.table + id * p.D;
    float s = 0.f;
    if (vec) {
      for (int k = lane * 4; k < p.D; k += 128) {
        const float4 a = *reinterpret_cast<const float4*>(f + k);
        const float4 b = __ldg(reinterpret_cast<const float4*>(e + k));
        s = fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, s))));
      }
    } else {
      for (int k = lane; k < p.D; k += 32) s = fmaf(f[k], __ldg(e + k), s);
    }
    s = warp_sum(s) + user_term;
    if (c == 0) l0 = s; else rank += (s > l0) ? 1 : 0;
    if (p.logits && lane == 0) p.logits[u * p.ldl + c] = s;
  }
  if (p.rank && lane == 0) p.rank[u] = rank;
}

__global__ void __launch_bounds__(256) add_user_term_kernel(float* logits, int ldl, int64_t U, int I, const float* feats_tail,
                                                           int ldf, const float* fake_table, const int64_t* label, int F) {
  const int64_t u = blockIdx.x;
  __shared__ float term;
  if (threadIdx.x < 32) {
    float s = 0.f;
    const int64_t lab = label[u];
    for (int c = threadIdx.x; c < F; c += 32) s = fmaf(feats_tail[u * ldf + c], __ldg(fake_table + lab * F + c), s);
    s = warp_sum(s);
    if (threadIdx.x == 0) term = s;
  }
  __syncthreads();
  const float t = term;
  for (int i = threadIdx.x; i < I; i += blockDim.x) logits[u * ldl + i] += t;
}

}  // namespace srfrd

using namespace srfrd;

extern "C" int srfrd_sample_candidates(const int64_t* offsets, const int* items, const int* users0, const int* target,
                                       int64_t U, int itemnum, int C, uint64_t seed, int64_t* cand, void* stream) {
  SRFRD_REQUIRE(offsets && items && users0 && target && cand, "sample_candidates: null pointer");
  SRFRD_REQUIRE(itemnum > 0 && C >= 1, "sample_candidates: bad itemnum / candidate count");
  if (U == 0) return 0;
  CandParams p;
  p.offsets = offsets; p.items = items; p.users0 = users0; p.target = target; p.itemnum = itemnum; p.C = C; p.U = U;
  p.seed = seed; p.cand = cand;
  const int64_t n = U * C;
  sample_candidates_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p);
  SRFRD_LAUNCH_CHECK();
  return 0;
}

extern "C" int srfrd_candidate_rank(const float* feats, int ldf, const float* item_table, int64_t n_rows, int D,
                                    const int64_t* cand, int64_t U, int C, const float* fake_table, int F,
                                    const int64_t* user_label, float* logits, int ldl, int* rank, int* err_flag,
                                    void* stream) {
  SRFRD_REQUIRE(feats && item_table && cand, "candidate_rank: null pointer");
  SRFRD_REQUIRE(logits || rank, "candidate_rank: no output");
  SRFRD_REQUIRE(D > 0 && C >= 1 && ldf >= D + (fake_table ? F : 0), "candidate_rank: bad widths (D=%d C=%d ldf=%d)", D, C, ldf);
  SRFRD_REQUIRE(!fake_table || user_label, "candidate_rank: SRFRN rows need the user labels");
  SRFRD_REQUIRE(!logits || ldl >= C, "candidate_rank: bad ldl");
  if (U == 0) return 0;
  RankParams p;
  p.feats = feats; p.ldf = ldf; p.table = item_table; p.D = D; p.n_rows = n_rows; p.cand = cand;
  p.fake_table = fake_table; p.F = F; p.user_label = user_label; p.U = U; p.C = C; p.logits = logits; p.ldl = ldl;
  p.rank = rank; p.err = err_flag;
  candidate_rank_kernel<<<(unsigned)((U + 7) / 8), 256, 0, (cudaStream_t)stream>>>(p);
  SRFRD_LAUNCH_CHECK();
  return 0;
}

extern "C" int srfrd_add_user_term(float* logits, int ldl, int64_t U, int I, const float* feats_tail, int ldf,
                                   const float* fake_table, const int64_t* label, int F, void* stream) {
  SRFRD_REQUIRE(logits && feats_tail && fake_table && label, "add_user_term: null pointer");
  if (U == 0 || I == 0) return 0;
  add_user_term_kernel<<<(unsigned)U, 256, 0, (cudaStream_t)stream>>>(logits, ldl, U, I, feats_tail, ldf, fake_table, label, F);
  SRFRD_LAUNCH_CHECK();
  return 0;
}
