/*
 * ds_tma.h -- bulk asynchronous copies (TMA, cp.async.bulk) + mbarrier helpers.
 *
 * The banded factorisation streams 8-row blocks of the H / L bands between HBM
 * and the shared-memory window.  Rows of a block are contiguous in both places,
 * so the non-tensor bulk form of TMA is enough: one elected thread issues
 * cp.async.bulk global->shared with an mbarrier that counts the bytes, the CTA
 * waits on the barrier's phase parity where it needs the rows.
 * Under emulation (g++) the copy is a memcpy and the waits are no-ops.
 */
#ifndef DS_TMA_H_
#define DS_TMA_H_
#include "ds_common.h"

namespace ds {

#if DS_CUDA
DS_FN uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

/* shared-memory accesses through 32-bit shared-window addresses (no generic->shared conversion
 * and no 64-bit address arithmetic in the inner loops) */
DS_FN double lds_f64(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
DS_FN dbl2 lds_v2f64(uint32_t a) {
  dbl2 v;
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
  return v;
}
DS_FN void sts_v2f64(uint32_t a, double x, double y) {
  asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(a), "d"(x), "d"(y) : "memory");
}
DS_FN int4 lds_v4s32(uint32_t a) {
  int4 v;
  asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}

DS_FN void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
DS_FN void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
/* order generic-proxy accesses (plain ld/st) before later async-proxy (TMA) accesses */
DS_FN void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

DS_FN void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
/* bytes: multiple of 16; dst/src 16-byte aligned */
DS_FN void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(__cvta_generic_to_global(src_gmem)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
/* L2 eviction policy for data that is streamed once per use (the H / L bands: 2 x 365 KB per CTA,
 * far more than the CTAs' share of L2): evict-first keeps the small re-read scratch (per-match and
 * per-facet sums, border rows, inputs) resident instead */
DS_FN uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
DS_FN uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
DS_FN uint64_t l2_policy_evict_normal() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
DS_FN void tma_load_1d_stream(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst_smem)),
      "l"(__cvta_generic_to_global(src_gmem)), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
      : "memory");
}
DS_FN void tma_store_1d_stream(void *dst_gmem, const void *src_smem, uint32_t bytes, uint64_t pol) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(
                   __cvta_generic_to_global(dst_gmem)),
               "r"(smem_u32(src_smem)), "r"(bytes), "l"(pol)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
/* plain store with the streaming (evict-first) cache operator */
DS_FN void st_stream(double *p, double v) { __stcs(p, v); }
DS_FN void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
/* shared -> global bulk store (bulk async-group completion) */
DS_FN void tma_store_1d(void *dst_gmem, const void *src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(__cvta_generic_to_global(dst_gmem)),
               "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
/* all committed bulk stores have finished READING their shared-memory source */
DS_FN void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
/* all committed bulk stores are complete (writes visible) */
DS_FN void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
DS_FN void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LAB_WAIT%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LAB_DONE%=;\n"
      "bra LAB_WAIT%=;\n"
      "LAB_DONE%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
/* one poll, no loop: 1 when the phase with this parity is complete */
DS_FN uint32_t mbar_try(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
#else
DS_FN void mbar_init(uint64_t *, int) {}
DS_FN void fence_mbar_init() {}
DS_FN void fence_proxy_async() {}
DS_FN void mbar_expect_tx(uint64_t *, uint32_t) {}
DS_FN void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *) { memcpy(dst, src, bytes); }
DS_FN uint64_t l2_policy_evict_first() { return 0; }
DS_FN uint64_t l2_policy_evict_last() { return 0; }
DS_FN uint64_t l2_policy_evict_normal() { return 0; }
DS_FN void st_stream(double *p, double v) { *p = v; }
DS_FN void tma_load_1d_stream(void *dst, const void *src, uint32_t bytes, uint64_t *, uint64_t) { memcpy(dst, src, bytes); }
DS_FN void tma_store_1d_stream(void *dst, const void *src, uint32_t bytes, uint64_t) { memcpy(dst, src, bytes); }
DS_FN void mbar_wait(uint64_t *, uint32_t) {}
DS_FN void fence_proxy_async_smem() {}
DS_FN void tma_store_1d(void *dst, const void *src, uint32_t bytes) { memcpy(dst, src, bytes); }
DS_FN void tma_store_wait_read() {}
DS_FN void tma_store_wait_all() {}
#endif

}  // namespace ds
#endif
