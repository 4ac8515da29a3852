/*
 * emu_sft.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * Compiles the kernel source of the SfT solve (defslam_b200/csrc/sft_core.h)
 * with plain g++ as a one-thread "team", so that the CPU-only test tier can
 * check the kernel's arithmetic, indexing and LM control flow against the
 * oracle on a box without a GPU.  It is NOT a fallback: nothing in the
 * defslam_b200 package loads this library, and the product library
 * (libdefslam_b200.so) returns DEFSLAM_ECUDA when no device is usable.
 * Built into tests/_emu/ by tests/emu/build.py.
 */
#define DS_EMULATE 1
#include <stdlib.h>

#include <map>
#include <memory>

#include "../../defslam_b200/csrc/ds_batch.h"

using namespace ds;

namespace {
struct EmuTemplate {
  PlanHost host;
  PlanView view;
};

int run(int nprob, const defslam_sft_problem *p, defslam_sft_result *r, int mode, double *H, double *b, double *chi) {
  std::map<const defslam_template_desc *, std::unique_ptr<EmuTemplate>> tmp;
  BatchMarshal bm;
  auto resolve = [&](const defslam_sft_problem &q, const PlanView **hv, const PlanView **dv) -> int {
    const EmuTemplate *t = (const EmuTemplate *)q.tmpl;
    if (!t) {
      if (!q.tmpl_desc) return DEFSLAM_EBADARG;
      auto it = tmp.find(q.tmpl_desc);
      if (it == tmp.end()) {
        std::unique_ptr<EmuTemplate> nt(new EmuTemplate);
        const int rc = nt->host.build(q.tmpl_desc);
        if (rc) return rc;
        nt->view = nt->host.host_view();
        it = tmp.emplace(q.tmpl_desc, std::move(nt)).first;
      }
      t = it->second.get();
    }
    *hv = &t->view;
    *dv = &t->view;
    return 0;
  };
  /* DEFSLAM_EMU_SMEM_LIMIT (doubles) lets a test force the placements the planner falls back to on
   * the device: border rows, then x/dx, in the global workspace */
  const char *lim = getenv("DEFSLAM_EMU_SMEM_LIMIT");
  if (const char *e = getenv("DEFSLAM_ROW_MODE")) bm.row_mode = atoi(e);
  int rc = bm.plan(nprob, p, mode, lim ? atoi(lim) : 1 << 28, resolve);
  if (rc) return rc;
  std::vector<uint8_t> in(bm.in_bytes + 16), out(bm.out_bytes + 16, 0);
  bm.pack_inputs(p, in.data());
  bm.bind(in.data(), out.data());
  const WorkspaceSizes z = bm.ws_sizes();
  std::vector<uint8_t> ws(workspace_bytes(z) + 64, 0);
  std::vector<double> smem(bm.smem_doubles + 8, 0.0);
  Team team;
  team.tid = 0;
  team.nthr = 1;
  for (int i = 0; i < nprob; i++) {
    /* poison the scratch so that stale-state bugs between problems show up */
    for (auto &v : smem) v = 1e300;
    if (bm.any_x_global) sft_run_problem<false>(team, bm.views[i], smem.data(), ws.data(), z, true, nullptr);
    else sft_run_problem<true>(team, bm.views[i], smem.data(), ws.data(), z, true, nullptr);
  }
  if (mode == MODE_NORMAL_EQ) {
    const ProbSlot &s = bm.slots[0];
    const size_t D = 3 * (size_t)s.n_nodes + 6;
    const ResultScalars *rs = (const ResultScalars *)(out.data() + s.out_off + s.o_res);
    if (rs->status) return rs->status;
    if (H) memcpy(H, out.data() + s.out_off + s.o_H, D * D * sizeof(double));
    if (b) memcpy(b, out.data() + s.out_off + s.o_b, D * sizeof(double));
    if (chi) *chi = rs->chi2_initial;
    return 0;
  }
  return bm.unpack(out.data(), r);
}
}  // namespace

extern "C" {
int emu_sft_solve_batched(int32_t nprob, const defslam_sft_problem *p, defslam_sft_result *r, int device) {
  (void)device;
  if (nprob < 0 || (nprob > 0 && (!p || !r))) return DEFSLAM_EBADARG;
  if (nprob == 0) return 0;
  return run(nprob, p, r, MODE_SOLVE, nullptr, nullptr, nullptr);
}
int emu_sft_solve(const defslam_sft_problem *p, defslam_sft_result *r) { return emu_sft_solve_batched(1, p, r, -1); }
int emu_sft_normal_equations(const defslam_sft_problem *p, double *H, double *b, double *chi2) {
  if (!p) return DEFSLAM_EBADARG;
  return run(1, p, nullptr, MODE_NORMAL_EQ, H, b, chi2);
}
int emu_template_create(const defslam_template_desc *d, int device, void **out) {
  (void)device;
  if (!out) return DEFSLAM_EBADARG;
  std::unique_ptr<EmuTemplate> t(new EmuTemplate);
  const int rc = t->host.build(d);
  if (rc) return rc;
  t->view = t->host.host_view();
  *out = t.release();
  return 0;
}
void emu_template_destroy(void *t) { delete (EmuTemplate *)t; }
int emu_plan_info(void *t, int32_t *out /* bw, ld, Dn_pad, Wr, n_blk, smem_doubles */) {
  const EmuTemplate *e = (const EmuTemplate *)t;
  out[0] = e->view.bw; out[1] = e->view.ld; out[2] = e->view.Dn_pad; out[3] = e->view.Wr; out[4] = e->view.n_blk;
  out[5] = CTX_DOUBLES + smem_layout(e->view.n_nodes, e->view.n_edges, e->view.Dn_pad, e->view.bwp, e->view.ld, e->view.Wr, e->view.ES, true).total;
  return 0;
}
}
