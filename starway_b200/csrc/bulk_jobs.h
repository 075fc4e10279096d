// starway_b200 — host-side preparation of a balanced bulk launch (shared by the CUDA backend and the
// CPU stand-in used by the host-logic tests).
#pragma once
#include <stdint.h>

#include "sw_device.h"

namespace swgpu {

constexpr uint64_t SW_BULK_MIN_SHARE = 32768;   // below this a CTA spends its time on launch latency

// Merges contiguous segments back into jobs (the engine cuts messages into segments) and sizes the
// grid so that every CTA gets an equal byte range.  Returns false when the launch does not qualify
// (more than SW_BULK_INLINE_JOBS jobs, or pieces that are not 16-byte aligned): the caller then uses
// the segment-list kernel.
inline bool bulk_build_jobs(const SwSeg* segs, uint32_t nseg, uint32_t max_ctas, uint32_t stage_bytes,
                            uint32_t nstages, SwBulkJobArgs* a, uint32_t* grid) {
  uint32_t n = 0;
  uint64_t total = 0;
  uint64_t job_len = 0;
  for (uint32_t i = 0; i < nseg; i++) {
    const SwSeg& g = segs[i];
    if (!g.len) continue;
    if ((g.src | g.dst | g.len) & 15) return false;
    if (n && a->src[n - 1] + job_len == g.src && a->dst[n - 1] + job_len == g.dst) {
      job_len += g.len;   // continues the current job
    } else {
      if (n == SW_BULK_INLINE_JOBS) return false;
      a->src[n] = g.src;
      a->dst[n] = g.dst;
      job_len = g.len;
      n++;
    }
    total += g.len;
    a->end[n - 1] = total;
  }
  if (!n || !max_ctas) return false;
  uint64_t share = (total + max_ctas - 1) / max_ctas;
  if (share < SW_BULK_MIN_SHARE) share = SW_BULK_MIN_SHARE;
  share = (share + 1023) & ~1023ull;
  a->njobs = n;
  a->stage_bytes = stage_bytes;
  a->nstages = nstages;
  a->pad = 0;
  a->share = share;
  *grid = (uint32_t)((total + share - 1) / share);
  return true;
}

}  // namespace swgpu
