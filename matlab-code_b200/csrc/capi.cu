// capi.cu - extern "C" boundary (include/aoadmm.h).  No C++ exception crosses the ABI: every entry point
// converts failures into an aoadmm_status and records a message retrievable with aoadmm_last_error().
#include <cstring>
#include <mutex>
#include <string>

#include "engine.h"

struct aoadmm_handle {
  aoadmm::Engine* eng = nullptr;
  std::string err;
};

namespace aoadmm {
void nccl_unique_id(uint8_t id[128]);
}

namespace {
std::string g_create_error;
std::mutex g_mu;

template <typename F>
int guard(aoadmm_handle* h, F&& fn) {
  try {
    fn();
    return AOADMM_OK;
  } catch (const aoadmm::CudaError& e) {
    if (h) h->err = e.what(); else { std::lock_guard<std::mutex> l(g_mu); g_create_error = e.what(); }
    cudaGetLastError();
    return e.code;
  } catch (const std::bad_alloc&) {
    if (h) h->err = "host allocation failed"; else { std::lock_guard<std::mutex> l(g_mu); g_create_error = "host allocation failed"; }
    return AOADMM_ERR_OOM;
  } catch (const std::exception& e) {
    if (h) h->err = e.what(); else { std::lock_guard<std::mutex> l(g_mu); g_create_error = e.what(); }
    return AOADMM_ERR_INVALID_ARG;
  }
}

struct DevBuf {
  double* p = nullptr;
  explicit DevBuf(size_t n) { AO_CUDA(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(double))); }
  ~DevBuf() { if (p) cudaFree(p); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
};

void select_device(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) throw aoadmm::CudaError(AOADMM_ERR_NO_DEVICE, "no CUDA device available");
  if (device < 0 || device >= n) throw aoadmm::CudaError(AOADMM_ERR_INVALID_ARG, "device ordinal out of range");
  AO_CUDA(cudaSetDevice(device));
}
}  // namespace

extern "C" {

int aoadmm_abi_version(void) { return AOADMM_ABI_VERSION; }

int aoadmm_device_count(int* count) {
  if (!count) return AOADMM_ERR_INVALID_ARG;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  *count = n;
  return AOADMM_OK;
}

int aoadmm_nccl_unique_id(uint8_t id[128]) {
  if (!id) return AOADMM_ERR_INVALID_ARG;
  return guard(nullptr, [&] { aoadmm::nccl_unique_id(id); });
}

int aoadmm_create(const aoadmm_problem* problem, const aoadmm_dist* dist, aoadmm_handle** out) {
  if (!out) return AOADMM_ERR_INVALID_ARG;
  *out = nullptr;
  aoadmm_handle* h = nullptr;
  int rc = guard(nullptr, [&] {
    h = new aoadmm_handle();
    h->eng = new aoadmm::Engine(problem, dist);
  });
  if (rc != AOADMM_OK) {
    if (h) {
      delete h->eng;
      delete h;
    }
    return rc;
  }
  *out = h;
  return AOADMM_OK;
}

int aoadmm_destroy(aoadmm_handle* h) {
  if (!h) return AOADMM_OK;
  delete h->eng;
  delete h;
  return AOADMM_OK;
}

const char* aoadmm_last_error(const aoadmm_handle* h) {
  if (h) return h->err.c_str();
  std::lock_guard<std::mutex> l(g_mu);
  static thread_local std::string copy;
  copy = g_create_error;
  return copy.c_str();
}

int aoadmm_set_state(aoadmm_handle* h, int32_t field, int32_t index, int32_t slice, const double* data, int64_t rows,
                     int64_t cols) {
  if (!h) return AOADMM_ERR_INVALID_ARG;
  return guard(h, [&] { h->eng->set_state(field, index, slice, data, rows, cols); });
}

int aoadmm_get_state(aoadmm_handle* h, int32_t field, int32_t index, int32_t slice, double* data, int64_t rows,
                     int64_t cols) {
  if (!h) return AOADMM_ERR_INVALID_ARG;
  return guard(h, [&] { h->eng->get_state(field, index, slice, data, rows, cols); });
}

int aoadmm_run(aoadmm_handle* h, const aoadmm_options* options, aoadmm_out* out) {
  if (!h) return AOADMM_ERR_INVALID_ARG;
  return guard(h, [&] { h->eng->run(options, out); });
}

int aoadmm_generate_cp_data(aoadmm_handle* h, int32_t object, const double* const* factors, double noise, uint64_t seed) {
  if (!h || !factors) return AOADMM_ERR_INVALID_ARG;
  return guard(h, [&] { h->eng->generate_cp_data(object, factors, noise, seed); });
}

int aoadmm_get_object_data(aoadmm_handle* h, int32_t object, double* out, int64_t n_elements) {
  if (!h || !out) return AOADMM_ERR_INVALID_ARG;
  return guard(h, [&] { h->eng->object_to_host(object, out, n_elements); });
}

int aoadmm_nvecs(aoadmm_handle* h, int32_t mode, int32_t slice, int32_t r, double* out, int64_t rows, double* info) {
  if (!h || !out || r < 1 || rows < 1) return AOADMM_ERR_INVALID_ARG;
  return guard(h, [&] { h->eng->nvecs_to_host(mode, slice, r, out, rows, info); });
}

int aoadmm_object_mttkrp(aoadmm_handle* h, int32_t object, int32_t pos, int32_t precision, double* out) {
  if (!h || !out || precision < 0 || precision > 1) return AOADMM_ERR_INVALID_ARG;
  return guard(h, [&] { h->eng->mttkrp_to_host(object, pos, out, precision); });
}

int aoadmm_time_mttkrp(aoadmm_handle* h, int32_t object, int32_t pos, int32_t reps, float* ms_out) {
  if (!h || !ms_out) return AOADMM_ERR_INVALID_ARG;
  return guard(h, [&] { *ms_out = h->eng->time_mttkrp(object, pos, reps); });
}

int aoadmm_launch_count(const aoadmm_handle* h, int64_t* count) {
  if (!h || !count) return AOADMM_ERR_INVALID_ARG;
  *count = h->eng->launches();
  return AOADMM_OK;
}

int aoadmm_phase_ms(const aoadmm_handle* h, double ms[3]) {
  if (!h || !ms) return AOADMM_ERR_INVALID_ARG;
  h->eng->phase_ms(ms);
  return AOADMM_OK;
}

int aoadmm_last_run_ms(const aoadmm_handle* h, double* ms) {
  if (!h || !ms) return AOADMM_ERR_INVALID_ARG;
  *ms = h->eng->last_run_ms();
  return AOADMM_OK;
}

int aoadmm_last_loop_ms(const aoadmm_handle* h, double* ms) {
  if (!h || !ms) return AOADMM_ERR_INVALID_ARG;
  *ms = h->eng->last_loop_ms();
  return AOADMM_OK;
}

// ---- operator-level entry points ---------------------------------------------------------------
int aoadmm_mttkrp(const double* X, int32_t order, const int64_t* dims, const double* const* factors, int32_t R,
                  int32_t n, double* out, int32_t device) {
  if (!X || !dims || !factors || !out || order < 2 || order > 8 || n < 1 || n > order || R < 1 || R > 256)
    return AOADMM_ERR_INVALID_ARG;
  return guard(nullptr, [&] {
    // a one-object problem driven through the engine so that exactly the production kernels run
    std::vector<int64_t> rows(dims, dims + order);
    std::vector<int32_t> rank(order, R), modes(order), lin(order, 0), constrained(order, 0);
    std::vector<aoadmm_constraint> cons(order);
    std::memset(cons.data(), 0, sizeof(aoadmm_constraint) * order);
    for (int d = 0; d < order; ++d) modes[d] = d + 1;
    aoadmm_object obj{};
    obj.model = AOADMM_MODEL_CP;
    obj.order = order;
    obj.modes = modes.data();
    obj.weight = 1.0;
    obj.znorm_const = 0.0;
    obj.data = X;
    obj.shard_offset = 0;
    obj.shard_extent = dims[order - 1];
    aoadmm_problem pb{};
    pb.nb_modes = order;
    pb.mode_rows = rows.data();
    pb.mode_rank = rank.data();
    pb.n_objects = 1;
    pb.objects = &obj;
    pb.lin_coupled_modes = lin.data();
    pb.n_couplings = 0;
    pb.constrained_modes = constrained.data();
    pb.constraints = cons.data();
    aoadmm_dist dist{};
    dist.rank = 0;
    dist.world_size = 1;
    dist.device = device;
    aoadmm::Engine eng(&pb, &dist);
    for (int d = 0; d < order; ++d) eng.set_state(AOADMM_FIELD_FAC, d + 1, 0, factors[d], dims[d], R);
    eng.mttkrp_to_host(1, n, out);
  });
}

int aoadmm_prox(const aoadmm_constraint* spec, const double* X, int64_t rows, int64_t cols, double rho, double* out,
                int32_t device) {
  if (!spec || !X || !out || rows < 0 || cols < 0 || cols > (1 << 20)) return AOADMM_ERR_INVALID_ARG;
  return guard(nullptr, [&] {
    select_device(device);
    const size_t n = (size_t)rows * cols;
    DevBuf dx(n), dout(n);
    const size_t sb = aoadmm::prox_scratch_bytes(spec->kind, rows, (int)cols);
    DevBuf scratch(sb / sizeof(double) + 1);
    AO_CUDA(cudaMemcpy(dx.p, X, n * sizeof(double), cudaMemcpyHostToDevice));
    if (spec->kind == AOADMM_CON_QUADRATIC) {
      if (spec->matrix_n != rows) throw aoadmm::CudaError(AOADMM_ERR_INVALID_ARG, "quadratic regularization: L must be rows x rows");
      aoadmm::QuadProx q;
      aoadmm::quad_prox_setup(q, spec->matrix, rows, spec->p0, (int)cols, 0);
      aoadmm::quad_prox_apply(q, dx.p, rows, dout.p, rows, (int)cols, nullptr, rho, 0, nullptr);
      AO_CUDA(cudaDeviceSynchronize());
      aoadmm::quad_prox_free(q);
      AO_CUDA(cudaMemcpy(out, dout.p, n * sizeof(double), cudaMemcpyDeviceToHost));
      return;
    }
    if (spec->kind == AOADMM_CON_TPARAFAC2)
      throw aoadmm::CudaError(AOADMM_ERR_UNSUPPORTED, "tPARAFAC2 acts on all PARAFAC2 slices at once: use the solver");
    aoadmm::prox_apply(spec->kind, spec->p0, spec->p1, dx.p, rows, dout.p, rows, rows, (int)cols, nullptr, rho,
                       sb ? scratch.p : nullptr, 0, nullptr);
    AO_CUDA(cudaDeviceSynchronize());
    AO_CUDA(cudaMemcpy(out, dout.p, n * sizeof(double), cudaMemcpyDeviceToHost));
  });
}

int aoadmm_chol_solve(const double* B, int32_t R, const double* A, int64_t rows, double* X, int32_t device) {
  if (!B || !A || !X || R < 1 || R > 256 || rows < 0) return AOADMM_ERR_INVALID_ARG;
  return guard(nullptr, [&] {
    select_device(device);
    const size_t RR = (size_t)R * R, n = (size_t)rows * R;
    DevBuf dB(RR), dC(RR), dBo(RR), dL(RR), dinv(R), drho(1), dA(n), dX(n);
    DevBuf ones(RR);
    aoadmm::InnerCtl* ctl = nullptr;
    AO_CUDA(cudaMalloc(&ctl, sizeof(aoadmm::InnerCtl)));
    AO_CUDA(cudaMemset(ctl, 0, sizeof(aoadmm::InnerCtl)));
    AO_CUDA(cudaMemcpy(dB.p, B, RR * sizeof(double), cudaMemcpyHostToDevice));
    AO_CUDA(cudaMemcpy(dA.p, A, n * sizeof(double), cudaMemcpyHostToDevice));
    aoadmm::PrepArgs a{};
    a.had[0] = dB.p;
    a.nhad = 1;
    a.R = R;
    a.weight = 1.0;
    a.rho_scale = 1.0;
    a.do_chol = 1;
    a.C = dC.p;
    a.B = dBo.p;
    a.L = dL.p;
    a.invdiag = dinv.p;
    a.rho = drho.p;
    a.ctl = ctl;
    aoadmm::prep_system(a, 0, nullptr);
    aoadmm::ls_solve(dA.p, rows, dL.p, dinv.p, dX.p, rows, rows, R, ctl, 0, nullptr);
    AO_CUDA(cudaDeviceSynchronize());
    aoadmm::InnerCtl hc;
    AO_CUDA(cudaMemcpy(&hc, ctl, sizeof(hc), cudaMemcpyDeviceToHost));
    cudaFree(ctl);
    if (hc.err != 0) throw aoadmm::CudaError(AOADMM_ERR_NOT_POSITIVE_DEFINITE, "matrix is not positive definite");
    AO_CUDA(cudaMemcpy(X, dX.p, n * sizeof(double), cudaMemcpyDeviceToHost));
  });
}

int aoadmm_gram(const double* F, int64_t rows, int32_t R, double* G, int32_t device) {
  if (!F || !G || R < 1 || R > 256 || rows < 0) return AOADMM_ERR_INVALID_ARG;
  return guard(nullptr, [&] {
    select_device(device);
    const size_t n = (size_t)rows * R;
    DevBuf dF(n), dG((size_t)R * R), ws(aoadmm::gram_ws_doubles(rows, R));
    AO_CUDA(cudaMemcpy(dF.p, F, n * sizeof(double), cudaMemcpyHostToDevice));
    aoadmm::gram(dF.p, rows, rows, R, dG.p, ws.p, 0, nullptr);
    AO_CUDA(cudaDeviceSynchronize());
    AO_CUDA(cudaMemcpy(G, dG.p, (size_t)R * R * sizeof(double), cudaMemcpyDeviceToHost));
  });
}

}  // extern "C"
