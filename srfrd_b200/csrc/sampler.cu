// On-device batch sampler: replaces WarpSampler_fr / sample_function_fr (utils.py:14-90), whose
// pure-Python per-sample loop plus multiprocessing.Queue cannot feed one B200 at batch 4096, and the seven
// host-to-device copies per step (trainer.py:29).  The leave-one-out training interactions live on the device in CSR
// form (offsets, items, labels, p_fake); one warp builds one sequence of the batch:
//   user   = eligible[hash(seed, step, b) mod n_eligible]      (users with > 1 train item, utils.py:25)
//   seq    = items[:-1] right-aligned, left-padded with 0; pos = items shifted by one (utils.py:40-48)
//   rsq/prs= the discriminator labels likewise {1 fake, 2 real};  nrs = 1 wherever pos != 0 (utils.py:52)
//   neg    = uniform item in 1..itemnum that is not in the user's train set, wherever pos != 0 (utils.py:27-31,50)
//   w_pos  = 1 - p_fake(pos) ("soft"), 1[p_fake(pos) < 0.5] ("mask") or 1[pos != 0] ("none"): row L's weights
// The step counter is read from device memory (the Adam step state), so a captured CUDA graph draws a new batch on
// every replay.
#include "common.cuh"
#include "srfrd_b200.h"

namespace srfrd {

struct SampleParams {
  const int64_t* offsets;   // (usernum + 1)
  const int* items;         // (nnz)
  const int8_t* labels;     // (nnz)
  const float* p_fake;      // (nnz) or null
  const int* eligible;      // (n_eligible) 0-based user rows
  int n_eligible, itemnum, B, L, policy;
  uint64_t seed;
  const float* step;        // device step counter or null
  int64_t *users, *seq, *rsq, *pos, *prs, *neg, *nrs;
  float* w_pos;             // (B, L) or null
};

__global__ void __launch_bounds__(256) sample_batch_kernel(SampleParams p) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= p.B) return;
  const uint64_t seed = mix_seed(p.seed, p.step);
  const int u0 = p.eligible[hash_u32(seed, 0x5A11u, (uint64_t)b) % (uint32_t)p.n_eligible];
  const int64_t a = p.offsets[u0];
  const int n = (int)(p.offsets[u0 + 1] - a);                   // train length >= 2
  if (lane == 0 && p.users) p.users[b] = u0 + 1;
  for (int t = lane; t < p.L; t += 32) {
    // slot t holds train index (n - 1) - (L - t) for seq and the next one for pos; valid if >= 0
    const int src = (n - 1) - (p.L - t);
    const bool valid = src >= 0;
    const int64_t o = (int64_t)b * p.L + t;
    int64_t sq = 0, rq = 0, ps = 0, pr = 0, ng = 0;
    float w = 0.f;
    if (valid) {
      sq = p.items[a + src]; rq = p.labels[a + src];
      ps = p.items[a + src + 1]; pr = p.labels[a + src + 1];
      const float pf = p.p_fake ? p.p_fake[a + src + 1] : 0.f;
      w = p.policy == 2 ? 1.f - pf : (p.policy == 1 ? (pf < 0.5f ? 1.f : 0.f) : 1.f);
      // rejection sampling against the user's own (short) item list
      // random_neq (utils.py:14-18) loops until the draw is outside the user's set; after 256 failed draws (a user who
      // rated almost the whole catalogue) walk upwards from the last draw to the next free id -- never a colliding id
      // while one exists (the reference would spin forever if none does; here the last draw is kept)
      int cand = 1;
      bool hit = true;
      for (uint32_t tries = 0; hit && tries < 256; ++tries) {
        cand = 1 + (int)(hash_u32(seed, 0x4E65u + tries, (uint64_t)o) % (uint32_t)p.itemnum);
        hit = false;
        for (int k = 0; k < n; ++k) hit |= p.items[a + k] == cand;
      }
      for (int walk = 0; hit && walk < p.itemnum; ++walk) {
        cand = cand % p.itemnum + 1;
        hit = false;
        for (int k = 0; k < n; ++k) hit |= p.items[a + k] == cand;
      }
      ng = cand;
    }
    p.seq[o] = sq; p.rsq[o] = rq; p.pos[o] = ps; p.prs[o] = pr; p.neg[o] = ng; p.nrs[o] = valid ? 1 : 0;
    if (p.w_pos) p.w_pos[o] = w;
  }
}

}  // namespace srfrd

using namespace srfrd;

extern "C" int srfrd_sample_batch(const int64_t* offsets, const int* items, const int8_t* labels, const float* p_fake,
                                  const int* eligible, int n_eligible, int itemnum, int B, int L, int policy,
                                  uint64_t seed, const float* step, int64_t* users, int64_t* seq, int64_t* rsq,
                                  int64_t* pos, int64_t* prs, int64_t* neg, int64_t* nrs, float* w_pos, void* stream) {
  SRFRD_REQUIRE(offsets && items && labels && eligible && seq && rsq && pos && prs && neg && nrs, "sample_batch: null pointer");
  SRFRD_REQUIRE(n_eligible > 0 && itemnum > 0 && L > 0, "sample_batch: empty data (n_eligible=%d itemnum=%d L=%d)", n_eligible, itemnum, L);
  SRFRD_REQUIRE(policy >= 0 && policy <= 2, "sample_batch: policy must be 0 (none), 1 (mask) or 2 (soft)");
  if (B == 0) return 0;
  SampleParams p;
  p.offsets = offsets; p.items = items; p.labels = labels; p.p_fake = p_fake; p.eligible = eligible;
  p.n_eligible = n_eligible; p.itemnum = itemnum; p.B = B; p.L = L; p.policy = policy; p.seed = seed; p.step = step;
  p.users = users; p.seq = seq; p.rsq = rsq; p.pos = pos; p.prs = prs; p.neg = neg; p.nrs = nrs; p.w_pos = w_pos;
  sample_batch_kernel<<<(B + 7) / 8, 256, 0, (cudaStream_t)stream>>>(p);
  SRFRD_LAUNCH_CHECK();
  return 0;
}
