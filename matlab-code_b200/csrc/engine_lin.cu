// engine_lin.cu - linear couplings (coupling types 1..5) of the AO-ADMM sweep.
//
// Reference: functions/cmtf_fun_AOADMM.m
//   :278-389   per-type system matrices before the coupled ADMM (B += rho/2*H*H' for type 2, +rho/2*I for 3/4,
//              B2 = rho/2*H'*H (+rho/2*I) for the Sylvester types 1/5)
//   :698-1075  ADMM_coupled_case1..5
//   :1118-1210 eval_res_ADMM_coupl_case1..5
// With  G_m(F) = {H F, F H, F, F, H F}  and  D_m(Delta) = {Delta, Delta, H Delta, Delta H, Delta H2}  (types 1..5)
// every case is the same loop:
//     F_m      = argmin  ... + rho_m/2 ||G_m(F) - D_m(Delta) + mu_m||^2           (Cholesky, or Sylvester for 1/5)
//     Delta    = argmin  sum_m rho_m ||G_m(F_m) + mu_m - D_m(Delta)||^2            (mean, or small normal equations)
//     mu_m    += G_m(F_m) - D_m(Delta)
// PARAFAC2 third modes inside a linearly coupled group (the kron/blkdiag variants, :283-297, :305-312 ...) are not
// covered: aoadmm_create reports AOADMM_ERR_UNSUPPORTED for them.
#include <algorithm>
#include <cmath>

#include "engine.h"

namespace aoadmm {

namespace {
void upload(DevMat& d, const double* src, int64_t rows, int64_t cols) {
  dev_alloc(d, rows, cols);
  if (rows * cols > 0) AO_CUDA(cudaMemcpy(d.p, src, d.bytes(), cudaMemcpyHostToDevice));
}
}  // namespace

void Engine::setup_linear_coupling(const aoadmm_problem* prob, int c) {
  const int ctype = coupling_type_[c - 1];
  if (ctype < 1 || ctype > 5) throw CudaError(1, "coupling_type must be in 0..5");
  if (prob->coupling_rows == nullptr || prob->coupling_cols == nullptr)
    throw CudaError(1, "coupling_rows / coupling_cols (shape of G.coupling_fac) are required for linear couplings");
  LinGroup& g = lin_groups_[c - 1];
  g.ctype = ctype;
  const int64_t dr = prob->coupling_rows[c - 1], dc = prob->coupling_cols[c - 1];
  if (dr <= 0 || dc <= 0) throw CudaError(1, "empty coupling_fac");
  dev_alloc(delta_[c - 1], dr, dc);
  dev_alloc(g.Dold, dr, dc);
  dev_alloc(g.Ddiff, dr, dc);
  for (auto& m : modes_) {
    if (m.coupling != c) continue;
    // example_script14: H C = Delta with C of a PARAFAC2 model; type 5 (H C = Delta H2) shares the (K*R)^2 system
    const bool par2c = (m.par2_role == 3 && (ctype == 1 || ctype == 5));
    // the first PARAFAC2 mode goes through the generic branches of :278-389 like a CP mode; the second one cannot be
    // coupled (:191); the third one has its own (K*R)^2 system for type 1 and per-slice Delta updates for types 4/5
    if (m.par2_role == 2) throw CudaError(1, "the second PARAFAC2 mode cannot be coupled");
    const bool par2row = (m.par2_role == 3 && (ctype == 2 || ctype == 3 || ctype == 4));
    if (m.par2_role == 3 && !par2c && !par2row) throw CudaError(1, "coupling_type must be in 0..5");
    const int i = m.id - 1;
    if (prob->trafo == nullptr || prob->trafo[i] == nullptr)
      throw CudaError(1, "coupl_trafo_matrices{" + std::to_string(m.id) + "} is required for coupling type " + std::to_string(ctype));
    lin_modes_.emplace_back();
    m.lin = (int)lin_modes_.size() - 1;
    g.modes.push_back(m.id);
    LinMode& lm = lin_modes_.back();
    lm.ctype = ctype;
    const int64_t hr = prob->trafo_rows[i], hc = prob->trafo_cols[i];
    upload(lm.H, prob->trafo[i], hr, hc);
    int64_t sr = 0, sc = 0;  // coupling space
    bool ok = true;
    switch (ctype) {
      case 1: ok = (hc == m.rows && hr == dr && dc == m.R); sr = hr; sc = m.R; break;          // H F = Delta
      case 2: ok = (hr == m.R && m.rows == dr && hc == dc); sr = m.rows; sc = hc; break;        // F H = Delta
      case 3: ok = (hr == m.rows && hc == dr && dc == m.R); sr = m.rows; sc = m.R; break;       // F = H Delta
      case 4: ok = (hc == m.R && m.rows == dr && hr == dc); sr = m.rows; sc = m.R; break;       // F = Delta H
      case 5: {
        if (prob->trafo2 == nullptr || prob->trafo2[i] == nullptr)
          throw CudaError(1, "coupl_trafo_matrices2{" + std::to_string(m.id) + "} is required for coupling type 5");
        const int64_t h2r = prob->trafo2_rows[i], h2c = prob->trafo2_cols[i];
        upload(lm.H2, prob->trafo2[i], h2r, h2c);
        ok = (hc == m.rows && hr == dr && h2r == dc && h2c == m.R);                             // H F = Delta H2
        sr = hr;
        sc = m.R;
        break;
      }
    }
    if (!ok) throw CudaError(1, "coupling " + std::to_string(c) + ": transformation / coupling_fac shapes of mode " +
                                    std::to_string(m.id) + " do not match coupling type " + std::to_string(ctype));
    dev_alloc(m.muD, sr, sc);
    for (DevMat* d : {&lm.S1, &lm.S2, &lm.S3}) dev_alloc(*d, sr, sc);
    dev_alloc(lm.tmpF, m.rows, m.R);
    dev_alloc(lm.tmpF2, m.rows, m.R);
    if (m.constrained) dev_alloc(lm.Zold, m.rows, m.R);
    if (ctype == 2) {
      dev_alloc(lm.HHt, m.R, m.R);
      launches_ += dgemm_small(0, 1, m.R, m.R, hc, 1.0, nullptr, lm.H.p, hr, lm.H.p, hr, 0.0, lm.HHt.p, m.R, st_, nullptr);
    }
    if (par2row) {
      if (g.par2row >= 0) throw CudaError(2, "at most one third PARAFAC2 mode per linearly coupled group");
      lm.par2row = true;
      g.par2row = (int)g.modes.size() - 1;
      if (ctype == 3) dev_alloc(lm.Hs, hr, hc);
      if (ctype == 4) {
        dev_alloc(lm.AAA, hr, hr);
        launches_ += dgemm_small(0, 1, hr, hr, hc, 1.0, nullptr, lm.H.p, hr, lm.H.p, hr, 0.0, lm.AAA.p, hr, st_, nullptr);
        if (hr > 64) throw CudaError(2, "coupling type 4 with a PARAFAC2 third mode: at most 64 columns in the coupling factor");
        AO_CUDA(cudaMalloc(&g.Minv, sizeof(double) * (size_t)m.rows * hr * hr));
      }
      if (ctype == 2) AO_CUDA(cudaMalloc(&g.wsum, sizeof(double) * (size_t)m.rows));
    } else if (par2c) {
      const int64_t n = m.rows * m.R;
      if (n * 8 > 40 * 1024) throw CudaError(2, "coupling type 1 with a PARAFAC2 third mode: K*R must be <= 5120");
      lm.par2c = true;
      dev_alloc(lm.HtH, m.rows, m.rows);
      launches_ += dgemm_small(1, 0, m.rows, m.rows, hr, 1.0, nullptr, lm.H.p, hr, lm.H.p, hr, 0.0, lm.HtH.p, m.rows, st_, nullptr);
      for (DevMat* d : {&lm.B2, &lm.B2L, &lm.B2B, &lm.B2C}) dev_alloc(*d, n, n);
      AO_CUDA(cudaMalloc(&lm.B2invdiag, sizeof(double) * n));
      AO_CUDA(cudaMalloc(&lm.Bsys3, sizeof(double) * m.rows * m.R * m.R));
      AO_CUDA(cudaMalloc(&lm.rho_stats, sizeof(double) * 2));
      lm.rho_A = lm.rho_stats;       // mean(rho{mm})  (:712)
      lm.rho_D = lm.rho_stats + 1;   // sum(rho{jj})   (:742)
      if (ctype == 5) {              // per-row Delta systems AA + rho_k * H2 H2' for the first size(Delta,1) slices (:1033-1051)
        if (g.par2row >= 0) throw CudaError(2, "at most one third PARAFAC2 mode per linearly coupled group");
        g.par2row = (int)g.modes.size() - 1;
        const int64_t q2 = lm.H2.rows;
        if (q2 > 64) throw CudaError(2, "coupling type 5 with a PARAFAC2 third mode: at most 64 columns in the coupling factor");
        if (dr > m.rows) throw CudaError(1, "coupling type 5 with a PARAFAC2 third mode: coupling_fac has more rows than slices");
        dev_alloc(lm.AAA, q2, q2);
        launches_ += dgemm_small(0, 1, q2, q2, lm.H2.cols, 1.0, nullptr, lm.H2.p, q2, lm.H2.p, q2, 0.0, lm.AAA.p, q2, st_, nullptr);
        AO_CUDA(cudaMalloc(&g.Minv, sizeof(double) * (size_t)dr * q2 * q2));
      }
    } else if (ctype == 1 || ctype == 5) {
      // eigen-decomposition H'H = U diag(lam) U' once: one-sided Jacobi on H (q x I)
      if (m.rows > 4096) throw CudaError(2, "coupling types 1/5 support at most 4096 rows in the coupled factor");
      dev_alloc(lm.U, m.rows, m.rows);
      dev_alloc(lm.VB, m.R, m.R);
      dev_alloc(lm.Bwork, m.R, m.R);
      AO_CUDA(cudaMalloc(&lm.lam, sizeof(double) * m.rows));
      AO_CUDA(cudaMalloc(&lm.muB, sizeof(double) * m.R));
      DevMat scratch;
      upload(scratch, prob->trafo[i], hr, hc);
      launches_ += jacobi_onesided(scratch.p, hr, (int)hc, lm.U.p, lm.lam, st_);
      std::vector<double> sig((size_t)hc);
      AO_CUDA(cudaMemcpyAsync(sig.data(), lm.lam, sizeof(double) * hc, cudaMemcpyDeviceToHost, st_));
      AO_CUDA(cudaStreamSynchronize(st_));
      for (auto& v : sig) v = v * v;
      AO_CUDA(cudaMemcpy(lm.lam, sig.data(), sizeof(double) * hc, cudaMemcpyHostToDevice));
      dev_free(scratch);
    }
  }
  if (g.modes.empty()) throw CudaError(1, "coupling id without modes");
  if ((int)g.modes.size() > 5) throw CudaError(2, "more than 5 modes in a linearly coupled group");
  AO_CUDA(cudaMalloc(&g.scal, sizeof(double) * 8));
  AO_CUDA(cudaMemset(g.scal, 0, sizeof(double) * 8));
  // small normal equations of the Delta update (:875-881, :941-962, :1028-1053)
  int64_t q = 0;
  const ModeState& m0 = mode(g.modes[0]);
  const LinMode& l0 = lin_modes_[m0.lin];
  if (ctype == 3) q = l0.H.cols;
  if (ctype == 4) q = l0.H.rows;
  if (ctype == 5) q = l0.H2.rows;
  if (q > 0) {
    if (q > 256) throw CudaError(2, "coupling types 3/4/5: the Delta normal equations support at most 256 unknowns per row");
    for (DevMat* d : {&g.AA, &g.AAL, &g.AAB, &g.AAC}) dev_alloc(*d, q, q);
    AO_CUDA(cudaMalloc(&g.AAinvdiag, sizeof(double) * q));
    if (ctype == 3) {
      dev_alloc(g.BB, m0.R, q);   // BB' (R x q)
      dev_alloc(g.Dt, m0.R, q);   // Delta' (R x q)
    } else {
      dev_alloc(g.BB, delta_[c - 1].rows, q);
    }
  }
}

// reduction jobs of one group (pointers are final once the per-mode buffers exist)
void Engine::lin_build_jobs(int c) {
  LinGroup& g = lin_groups_[c - 1];
  std::vector<RedJob> jobs;
  auto add = [&](int kind, const double* a, const double* b, int64_t rows, int64_t cols) {
    RedJob j{};
    j.kind = kind;
    j.cols = (int)cols;
    j.rows = rows;
    j.lda = rows;
    j.ldb = rows;
    j.a = a;
    j.b = b;
    jobs.push_back(j);
    return (int)jobs.size() - 1;
  };
  g.fin.nmodes = (int)g.modes.size();
  for (size_t t = 0; t < g.modes.size(); ++t) {
    ModeState& m = mode(g.modes[t]);
    LinMode& lm = lin_modes_[m.lin];
    if (!lm.par2c) lm.rho_A = lm.rho_D = m.rho;   // (m.rho exists only now: per-mode buffers are allocated after the groups)
    LinFinMode& f = g.fin.m[t];
    f.i_pr_num = add(RED_DIFF2, lm.S1.p, lm.S2.p, lm.S1.rows, lm.S1.cols);
    const int i_f = add(RED_NORM2, m.fac.p, nullptr, m.rows, m.R);
    // :1124, :1143 divide by ||G(F)||, :1162, :1181, :1200 by ||F||
    f.i_pr_den = (g.ctype == 1 || g.ctype == 2) ? add(RED_NORM2, lm.S1.p, nullptr, lm.S1.rows, lm.S1.cols) : i_f;
    f.i_du_num = add(RED_NORM2, lm.S3.p, nullptr, lm.S3.rows, lm.S3.cols);
    f.i_mu = add(RED_NORM2, m.muD.p, nullptr, m.muD.rows, m.muD.cols);
    f.constrained = m.constrained ? 1 : 0;
    f.i_fn = i_f;
    f.i_fz = f.i_zz = f.i_muz = 0;
    if (m.constrained) {
      f.i_fz = add(RED_DIFF2, m.fac.p, m.Z.p, m.rows, m.R);
      f.i_zz = add(RED_DIFF2, m.Z.p, lm.Zold.p, m.rows, m.R);
      f.i_muz = add(RED_NORM2, m.muZ.p, nullptr, m.rows, m.R);
    }
  }
  g.njobs = (int)jobs.size();
  AO_CUDA(cudaMalloc(&g.jobs_dev, sizeof(RedJob) * g.njobs));
  AO_CUDA(cudaMemcpy(g.jobs_dev, jobs.data(), sizeof(RedJob) * g.njobs, cudaMemcpyHostToDevice));
  AO_CUDA(cudaMalloc(&g.red, sizeof(double) * g.njobs));
  AO_CUDA(cudaMalloc(&g.red_partials, sizeof(double) * reduce_ws_doubles(g.njobs)));
}

void Engine::free_linear_coupling() {
  for (auto& lm : lin_modes_) {
    for (DevMat* d : {&lm.H, &lm.H2, &lm.HHt, &lm.tmpF, &lm.tmpF2, &lm.S1, &lm.S2, &lm.S3, &lm.Zold, &lm.U, &lm.VB, &lm.Bwork})
      dev_free(*d);
    if (lm.lam) cudaFree(lm.lam);
    if (lm.muB) cudaFree(lm.muB);
    for (DevMat* d : {&lm.HtH, &lm.B2, &lm.B2L, &lm.B2B, &lm.B2C, &lm.Hs, &lm.AAA}) dev_free(*d);
    for (void* q : {(void*)lm.B2invdiag, (void*)lm.Bsys3, (void*)lm.rho_stats})
      if (q) cudaFree(q);
  }
  for (auto& g : lin_groups_) {
    for (DevMat* d : {&g.Dold, &g.Ddiff, &g.AA, &g.AAL, &g.AAB, &g.AAC, &g.BB, &g.Dt}) dev_free(*d);
    for (void* p : {(void*)g.AAinvdiag, (void*)g.scal, (void*)g.jobs_dev, (void*)g.red, (void*)g.red_partials,
                    (void*)g.wsum, (void*)g.Minv})
      if (p) cudaFree(p);
  }
}

// out (coupling space) = G_m(F)
void Engine::lin_G(const LinMode& lm, const ModeState& m, const double* F, double* out, const int* skip) {
  switch (lm.ctype) {
    case 1:
    case 5:
      launches_ += dgemm_small(0, 0, lm.H.rows, m.R, m.rows, 1.0, nullptr, lm.H.p, lm.H.rows, F, m.rows, 0.0, out, lm.H.rows, st_, skip);
      break;
    case 2:
      launches_ += dgemm_small(0, 0, m.rows, lm.H.cols, m.R, 1.0, nullptr, F, m.rows, lm.H.p, lm.H.rows, 0.0, out, m.rows, st_, skip);
      break;
    default: {
      LinTerm t{F, 1.0, nullptr};
      launches_ += lincomb(out, m.rows * m.R, &t, 1, st_, skip);
    }
  }
}

// out (coupling space) = D_m(Delta);  Dshape carries the shape of Delta
void Engine::lin_D(const LinMode& lm, const ModeState& m, const double* Delta, const DevMat& Dshape, double* out,
                   const int* skip) {
  switch (lm.ctype) {
    case 3:
      launches_ += dgemm_small(0, 0, m.rows, m.R, lm.H.cols, 1.0, nullptr, lm.H.p, lm.H.rows, Delta, Dshape.rows, 0.0, out, m.rows, st_, skip);
      break;
    case 4:
      launches_ += dgemm_small(0, 0, m.rows, m.R, lm.H.rows, 1.0, nullptr, Delta, Dshape.rows, lm.H.p, lm.H.rows, 0.0, out, m.rows, st_, skip);
      break;
    case 5:
      launches_ += dgemm_small(0, 0, Dshape.rows, m.R, lm.H2.rows, 1.0, nullptr, Delta, Dshape.rows, lm.H2.p, lm.H2.rows, 0.0, out, Dshape.rows, st_, skip);
      break;
    default: {
      LinTerm t{Delta, 1.0, nullptr};
      launches_ += lincomb(out, Dshape.rows * Dshape.cols, &t, 1, st_, skip);
    }
  }
}

// out (factor shape) = adjoint of G_m applied to Y (coupling space)
void Engine::lin_Gt(const LinMode& lm, const ModeState& m, const double* Y, double* out, const int* skip) {
  switch (lm.ctype) {
    case 1:
    case 5:
      launches_ += dgemm_small(1, 0, m.rows, m.R, lm.H.rows, 1.0, nullptr, lm.H.p, lm.H.rows, Y, lm.H.rows, 0.0, out, m.rows, st_, skip);
      break;
    case 2:
      launches_ += dgemm_small(0, 1, m.rows, m.R, lm.H.cols, 1.0, nullptr, Y, m.rows, lm.H.p, lm.H.rows, 0.0, out, m.rows, st_, skip);
      break;
    default: {
      LinTerm t{Y, 1.0, nullptr};
      launches_ += lincomb(out, m.rows * m.R, &t, 1, st_, skip);
    }
  }
}

// per outer iteration, after the per-mode precompute (rho_m is fixed during the inner loop)
void Engine::lin_prepare_group(int c) {
  LinGroup& g = lin_groups_[c - 1];
  const int n = (int)g.modes.size();
  InnerCtl* ctl = mode(g.modes[0]).ctl;
  for (int i = 0; i < n; ++i) {
    ModeState& m = mode(g.modes[i]);
    LinMode& lm = lin_modes_[m.lin];
    if (!lm.par2c) continue;
    Par2State& s = par2_[m.par2];
    launches_ += par2_rho_stats(s.rho3, s.K, lm.rho_stats, st_);
    launches_ += par2_assemble_B2(lm.Bsys3, lm.HtH.p, lm.rho_stats, m.constrained ? 1 : 0, s.K, s.R, lm.B2.p, st_);
    PrepArgs a{};   // L{m} = chol(B2{m},'lower')  (:296)
    a.nhad = 1;
    a.had[0] = lm.B2.p;
    a.R = (int)lm.B2.rows;
    a.weight = 1.0;
    a.rho_scale = 1.0;
    a.do_chol = 1;
    a.C = lm.B2C.p;
    a.B = lm.B2B.p;
    a.L = lm.B2L.p;
    a.invdiag = lm.B2invdiag;
    a.rho = g.scal + 3;
    a.ctl = ctl;
    launches_ += prep_system(a, st_, nullptr);
  }
  if ((g.ctype == 1 || g.ctype == 2) && g.par2row < 0) {
    LinTerm t[5];
    for (int i = 0; i < n; ++i) t[i] = LinTerm{nullptr, 1.0, lin_modes_[mode(g.modes[i]).lin].rho_D};
    launches_ += sum_recip(g.scal, t, n, st_);
  }
  if (g.ctype == 1 || g.ctype == 5) {
    for (int i = 0; i < n; ++i) {  // B = VB diag(muB) VB'
      ModeState& m = mode(g.modes[i]);
      LinMode& lm = lin_modes_[m.lin];
      if (lm.par2c) continue;
      AO_CUDA(cudaMemcpyAsync(lm.Bwork.p, m.B.p, m.B.bytes(), cudaMemcpyDeviceToDevice, st_));
      launches_ += jacobi_onesided(lm.Bwork.p, m.R, m.R, lm.VB.p, lm.muB, st_);
    }
  }
  if (g.ctype >= 3) {
    // AA = sum_j rho_j H_j'H_j (3) | sum_j rho_j H_j H_j' (4) | sum_j rhoC H2_j H2_j' (5, rhoC = rho of the LAST mode, :1032)
    // rhoC = mean(rho) of the LAST mode of the group (:1032; a scalar rho is its own mean)
    const double* rhoC = lin_modes_[mode(g.modes[n - 1]).lin].par2c ? lin_modes_[mode(g.modes[n - 1]).lin].rho_stats
                                                                      : mode(g.modes[n - 1]).rho;
    bool first = true;
    for (int i = 0; i < n; ++i) {
      ModeState& m = mode(g.modes[i]);
      LinMode& lm = lin_modes_[m.lin];
      const long long q = g.AA.rows;
      if (lm.par2row && g.ctype == 4) continue;   // its rho_k * H H' enters the per-row systems below (:944-951)
      if (lm.par2c && g.ctype == 5) continue;     // likewise rho_k * H2 H2' (:1033-1040)
      const double beta = first ? 0.0 : 1.0;
      first = false;
      if (g.ctype == 3 && lm.par2row) {            // H' diag(rho) H (:878 with a vector rho)
        launches_ += par2_rows_scale(lm.Hs.p, lm.H.p, m.rho_rows, lm.H.rows, (int)lm.H.cols, st_, nullptr);
        launches_ += dgemm_small(1, 0, q, q, lm.H.rows, 1.0, nullptr, lm.H.p, lm.H.rows, lm.Hs.p, lm.H.rows, beta, g.AA.p, q, st_, nullptr);
      } else if (g.ctype == 3)
        launches_ += dgemm_small(1, 0, q, q, lm.H.rows, 1.0, m.rho, lm.H.p, lm.H.rows, lm.H.p, lm.H.rows, beta, g.AA.p, q, st_, nullptr);
      else if (g.ctype == 4)
        launches_ += dgemm_small(0, 1, q, q, lm.H.cols, 1.0, m.rho, lm.H.p, lm.H.rows, lm.H.p, lm.H.rows, beta, g.AA.p, q, st_, nullptr);
      else
        launches_ += dgemm_small(0, 1, q, q, lm.H2.cols, 1.0, rhoC, lm.H2.p, lm.H2.rows, lm.H2.p, lm.H2.rows, beta, g.AA.p, q, st_, nullptr);
    }
    if (g.ctype == 4 && g.par2row >= 0) {
      // Delta(k,:) = BB(k,:) / (AA + rho_k * H H')  (:957-960): K small inverses per outer iteration
      ModeState& mp = mode(g.modes[g.par2row]);
      LinMode& lp = lin_modes_[mp.lin];
      if (first) AO_CUDA(cudaMemsetAsync(g.AA.p, 0, g.AA.bytes(), st_));   // no other mode in the group
      launches_ += par2_rowsys_inverse(g.AA.p, lp.AAA.p, mp.rho_rows, (int)mp.rows, (int)g.AA.rows, g.Minv, ctl, st_);
      return;
    }
    if (g.ctype == 5 && g.par2row >= 0) {
      ModeState& mp = mode(g.modes[g.par2row]);
      LinMode& lp = lin_modes_[mp.lin];
      if (first) AO_CUDA(cudaMemsetAsync(g.AA.p, 0, g.AA.bytes(), st_));
      launches_ += par2_rowsys_inverse(g.AA.p, lp.AAA.p, mp.rho_rows, (int)delta_[c - 1].rows, (int)g.AA.rows, g.Minv, ctl, st_);
      return;
    }
    PrepArgs a{};
    a.nhad = 1;
    a.had[0] = g.AA.p;
    a.R = (int)g.AA.rows;
    a.weight = 1.0;
    a.rho_scale = 1.0;
    a.do_chol = 1;
    a.C = g.AAC.p;
    a.B = g.AAB.p;
    a.L = g.AAL.p;
    a.invdiag = g.AAinvdiag;
    a.rho = g.scal + 2;
    a.ctl = ctl;
    launches_ += prep_system(a, st_, nullptr);
  }
}

void Engine::run_admm_linear(int c, std::vector<ModeState*>& group, const aoadmm_options& opt) {
  LinGroup& g = lin_groups_[c - 1];
  DevMat& D = delta_[c - 1];
  const int n = (int)group.size();
  InnerCtl* ctl = group[0]->ctl;
  const int* skip = &ctl->done;
  InnerTol tol{opt.innerRelPrTol_coupl, opt.innerRelDualTol_coupl, opt.innerRelPrTol_constr, opt.innerRelDualTol_constr};
  for (ModeState* mp : group) ++mp->version;
  const long long nD = D.rows * D.cols;
  for (int it = 0; it < opt.MaxInnerIters; ++it) {
    // ---- factor updates (:707-735, :782-805, :847-870, :913-936, :995-1023)
    for (ModeState* mp : group) {
      ModeState& m = *mp;
      LinMode& lm = lin_modes_[m.lin];
      const long long nF = m.rows * m.R, nS = lm.S1.rows * lm.S1.cols;
      lin_D(lm, m, D.p, D, lm.S2.p, skip);
      {
        LinTerm t[2] = {{lm.S2.p, 1.0, nullptr}, {m.muD.p, -1.0, nullptr}};
        launches_ += lincomb(lm.S2.p, nS, t, 2, st_, skip);
      }
      lin_Gt(lm, m, lm.S2.p, lm.tmpF.p, skip);
      if (lm.par2row) {
        // row k: A_inner = A{m}{k}' + rho_k/2 * (coupling term [+ Z - mu_Z]);  F(k,:) = (A_inner/L_k')/L_k  (:786-790 ...)
        launches_ += par2_rows_ainner(lm.tmpF2.p, m.A.p, lm.tmpF.p, m.constrained ? m.Z.p : nullptr, m.muZ.p, m.rho_rows,
                                      m.rows, m.R, st_, skip);
        launches_ += par2_rows_apply(m.fac.p, lm.tmpF2.p, m.Binv_rows, m.rows, m.R, st_, skip);
        continue;
      }
      {
        LinTerm t[4] = {{m.A.p, 1.0, nullptr}, {lm.tmpF.p, 0.5, lm.rho_A}, {m.Z.p, 0.5, lm.rho_A}, {m.muZ.p, -0.5, lm.rho_A}};
        launches_ += lincomb(lm.tmpF.p, nF, t, m.constrained ? 4 : 2, st_, skip);
      }
      if (lm.par2c) {
        // Gfacmm_vec = L'\(L\A_inner) on the (K*R) system, rows of C in the order k*R + r (:721-722)
        launches_ += par2_chol_solve_vec(lm.B2L.p, (int)m.rows, m.R, lm.tmpF.p, m.fac.p, st_, skip);
      } else if (g.ctype == 1 || g.ctype == 5) {
        // sylvester(B2,B,A_inner) (:728, :1016) with B2 = rho/2 (H'H [+ I]) = U (rho/2 (lam [+1])) U', B = VB muB VB'
        launches_ += dgemm_small(1, 0, m.rows, m.R, m.rows, 1.0, nullptr, lm.U.p, m.rows, lm.tmpF.p, m.rows, 0.0, lm.tmpF2.p, m.rows, st_, skip);
        launches_ += dgemm_small(0, 0, m.rows, m.R, m.R, 1.0, nullptr, lm.tmpF2.p, m.rows, lm.VB.p, m.R, 0.0, lm.tmpF.p, m.rows, st_, skip);
        launches_ += sylvester_scale(lm.tmpF2.p, lm.tmpF.p, m.rows, m.R, lm.lam, m.constrained ? 1.0 : 0.0, lm.muB, m.rho, st_, skip);
        launches_ += dgemm_small(0, 0, m.rows, m.R, m.rows, 1.0, nullptr, lm.U.p, m.rows, lm.tmpF2.p, m.rows, 0.0, lm.tmpF.p, m.rows, st_, skip);
        launches_ += dgemm_small(0, 1, m.rows, m.R, m.R, 1.0, nullptr, lm.tmpF.p, m.rows, lm.VB.p, m.R, 0.0, m.fac.p, m.rows, st_, skip);
      } else {
        // F = (A_inner/L')/L = A_inner * inv(B)
        launches_ += dgemm_small(0, 0, m.rows, m.R, m.R, 1.0, nullptr, lm.tmpF.p, m.rows, m.Binv.p, m.R, 0.0, m.fac.p, m.rows, st_, skip);
      }
    }
    // ---- Delta update
    AO_CUDA(cudaMemcpyAsync(g.Dold.p, D.p, D.bytes(), cudaMemcpyDeviceToDevice, st_));
    const double* rhoC = lin_modes_[group[n - 1]->lin].par2c ? lin_modes_[group[n - 1]->lin].rho_stats : group[n - 1]->rho;
    for (int i = 0; i < n; ++i) {
      ModeState& m = *group[i];
      LinMode& lm = lin_modes_[m.lin];
      const long long nS = lm.S1.rows * lm.S1.cols;
      lin_G(lm, m, m.fac.p, lm.S1.p, skip);
      if (g.ctype == 2 && g.par2row >= 0) {   // :808-815 with a vector rho: row-wise weighted mean
        launches_ += par2_rows_weighted_accum(D.p, g.wsum, lm.S1.p, m.muD.p, m.rho, lm.par2row ? m.rho_rows : nullptr, D.rows,
                                              (int)D.cols, i == 0 ? 1 : 0, st_, skip);
      } else if (g.ctype == 1 || g.ctype == 2) {  // :738-749, :808-815 rho-weighted mean
        if (i == 0) {
          LinTerm t[2] = {{lm.S1.p, 1.0, lm.rho_D}, {m.muD.p, 1.0, lm.rho_D}};
          launches_ += lincomb(D.p, nD, t, 2, st_, skip);
        } else {
          LinTerm t[3] = {{D.p, 1.0, nullptr}, {lm.S1.p, 1.0, lm.rho_D}, {m.muD.p, 1.0, lm.rho_D}};
          launches_ += lincomb(D.p, nD, t, 3, st_, skip);
        }
      } else {
        LinTerm t[2] = {{lm.S1.p, 1.0, nullptr}, {m.muD.p, 1.0, nullptr}};
        launches_ += lincomb(lm.S3.p, nS, t, 2, st_, skip);   // G(F) + mu
        const double beta = (i == 0) ? 0.0 : 1.0;
        const double* rs = m.rho;   // scalar weight; a third PARAFAC2 mode scales its rows by rho_k first
        if (lm.par2row) {
          launches_ += par2_rows_scale(lm.S3.p, lm.S3.p, m.rho_rows, m.rows, m.R, st_, skip);
          rs = nullptr;
        }
        if (g.ctype == 3)       // BB' (R x q) += rho (F+mu)' H            (:879)
          launches_ += dgemm_small(1, 0, m.R, lm.H.cols, m.rows, 1.0, rs, lm.S3.p, m.rows, lm.H.p, lm.H.rows, beta, g.BB.p, g.BB.rows, st_, skip);
        else if (g.ctype == 4)  // BB (I x q) += rho (F+mu) H'              (:955)
          launches_ += dgemm_small(0, 1, m.rows, lm.H.rows, m.R, 1.0, rs, lm.S3.p, m.rows, lm.H.p, lm.H.rows, beta, g.BB.p, g.BB.rows, st_, skip);
        else                    // BB (q1 x q2) += rhoC (H F + mu) H2'      (:1046)
          launches_ += dgemm_small(0, 1, lm.S3.rows, lm.H2.rows, m.R, 1.0, rhoC, lm.S3.p, lm.S3.rows, lm.H2.p, lm.H2.rows, beta, g.BB.p, g.BB.rows, st_, skip);
      }
    }
    if (g.ctype == 2 && g.par2row >= 0) {
      launches_ += par2_rows_divide(D.p, g.wsum, D.rows, (int)D.cols, st_, skip);
    } else if ((g.ctype == 4 || g.ctype == 5) && g.par2row >= 0) {   // Delta(k,:) = BB(k,:) / (AA + rho_k H H')  (:957-960, :1048-1051)
      launches_ += par2_rows_apply(D.p, g.BB.p, g.Minv, D.rows, (int)D.cols, st_, skip);
    } else if (g.ctype == 1 || g.ctype == 2) {
      LinTerm t{D.p, 1.0, g.scal + 1};
      launches_ += lincomb(D.p, nD, &t, 1, st_, skip);
    } else if (g.ctype == 3) {   // Delta = AA\BB  (:881)  <=>  Delta' = BB' inv(AA)
      launches_ += ls_solve(g.BB.p, g.BB.rows, g.AAL.p, g.AAinvdiag, g.Dt.p, g.Dt.rows, g.BB.rows, (int)g.AA.rows, ctl, st_, skip);
      launches_ += transpose_small(g.Dt.p, g.Dt.rows, g.Dt.cols, D.p, st_, skip);
    } else {                     // Delta = BB/AA  (:962, :1053)
      launches_ += ls_solve(g.BB.p, g.BB.rows, g.AAL.p, g.AAinvdiag, D.p, D.rows, g.BB.rows, (int)g.AA.rows, ctl, st_, skip);
    }
    {
      LinTerm t[2] = {{D.p, 1.0, nullptr}, {g.Dold.p, -1.0, nullptr}};
      launches_ += lincomb(g.Ddiff.p, nD, t, 2, st_, skip);
    }
    // ---- duals, constraints (:752-757 ...), residual operands
    for (ModeState* mp : group) {
      ModeState& m = *mp;
      LinMode& lm = lin_modes_[m.lin];
      const long long nF = m.rows * m.R, nS = lm.S1.rows * lm.S1.cols;
      lin_D(lm, m, D.p, D, lm.S2.p, skip);          // S1 = G(F) from the Delta step, S2 = D(Delta)
      lin_D(lm, m, g.Ddiff.p, D, lm.S3.p, skip);    // S3 = D(Delta - Delta_old)
      {
        LinTerm t[3] = {{m.muD.p, 1.0, nullptr}, {lm.S1.p, 1.0, nullptr}, {lm.S2.p, -1.0, nullptr}};
        launches_ += lincomb(m.muD.p, nS, t, 3, st_, skip);
      }
      if (m.constrained) {  // update_constraint (:1420-1429)
        AO_CUDA(cudaMemcpyAsync(lm.Zold.p, m.Z.p, m.Z.bytes(), cudaMemcpyDeviceToDevice, st_));
        LinTerm t[2] = {{m.fac.p, 1.0, nullptr}, {m.muZ.p, 1.0, nullptr}};
        launches_ += lincomb(lm.tmpF.p, nF, t, 2, st_, skip);
        launches_ += apply_prox(m, lm.tmpF.p, m.rows, m.Z.p, m.rows, m.rows, m.R, m.rho, skip);
        LinTerm u[3] = {{m.muZ.p, 1.0, nullptr}, {m.fac.p, 1.0, nullptr}, {m.Z.p, -1.0, nullptr}};
        launches_ += lincomb(m.muZ.p, nF, u, 3, st_, skip);
      }
    }
    launches_ += reduce_jobs(g.jobs_dev, g.njobs, g.red, g.red_partials, admm_counter_, st_, skip);
    launches_ += lin_finalize(g.fin, g.red, tol, ctl, st_);
  }
}

}  // namespace aoadmm
