// aoadmm_nvecs_mex.cpp - MEX gateway of the device-side nvecs initialisation.
//
//   U = aoadmm_nvecs_mex(X, i, r)
//     X : dense N-d double array (2 <= N <= 8), the .data of the tensor (or the matrix) that contains the mode
//     i : position of the mode inside X (1-based)
//     r : number of vectors
// replaces the body of functions/cmtf_nvecs.m:33-58 for one object ( A = tenmat(X,i); Y = A*A'; eigs(Y,r,'LM') ):
// the tensor is copied to the GPU once, the unfolding is never materialised, Y and its leading eigenvectors are
// computed on the device (include/aoadmm.h: aoadmm_nvecs).  Build:
//   mex -R2018a -I<repo>/include aoadmm_nvecs_mex.cpp -L<repo>/matlab-code_b200/aoadmm_b200 -laoadmm_b200
#include <cstdint>
#include <cstring>
#include <vector>

#include "aoadmm.h"
#include "mex.h"

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs != 3 || nlhs > 1) mexErrMsgIdAndTxt("aoadmm:invalidArg", "usage: U = aoadmm_nvecs_mex(X, i, r)");
  const mxArray* X = prhs[0];
  if (!mxIsDouble(X) || mxIsComplex(X) || mxIsSparse(X))
    mexErrMsgIdAndTxt("aoadmm:unsupported", "X must be a dense real double array");
  const int order = (int)mxGetNumberOfDimensions(X);
  const mwSize* dims = mxGetDimensions(X);
  const int pos = (int)mxGetScalar(prhs[1]), r = (int)mxGetScalar(prhs[2]);
  if (order < 2 || order > 8 || pos < 1 || pos > order) mexErrMsgIdAndTxt("aoadmm:invalidArg", "mode position out of range");

  std::vector<int64_t> rows(order);
  std::vector<int32_t> rank(order, r), modes(order), lin(order, 0), constrained(order, 0);
  std::vector<aoadmm_constraint> cons(order);
  std::memset(cons.data(), 0, sizeof(aoadmm_constraint) * order);
  for (int d = 0; d < order; ++d) {
    rows[d] = (int64_t)dims[d];
    modes[d] = d + 1;
  }
  aoadmm_object obj;
  std::memset(&obj, 0, sizeof(obj));
  obj.model = AOADMM_MODEL_CP;
  obj.order = order;
  obj.modes = modes.data();
  obj.weight = 1.0;
  obj.znorm_const = 0.0;
  obj.data = mxGetPr(X);
  obj.shard_offset = 0;
  obj.shard_extent = rows[order - 1];
  aoadmm_problem pb;
  std::memset(&pb, 0, sizeof(pb));
  pb.nb_modes = order;
  pb.mode_rows = rows.data();
  pb.mode_rank = rank.data();
  pb.n_objects = 1;
  pb.objects = &obj;
  pb.lin_coupled_modes = lin.data();
  pb.n_couplings = 0;
  pb.constrained_modes = constrained.data();
  pb.constraints = cons.data();
  aoadmm_dist dist;
  std::memset(&dist, 0, sizeof(dist));
  dist.world_size = 1;

  aoadmm_handle* h = nullptr;
  int st = aoadmm_create(&pb, &dist, &h);
  if (st != AOADMM_OK) mexErrMsgIdAndTxt("aoadmm:create", "%s", aoadmm_last_error(h));
  plhs[0] = mxCreateDoubleMatrix((mwSize)rows[pos - 1], (mwSize)r, mxREAL);
  st = aoadmm_nvecs(h, pos, 0, r, mxGetPr(plhs[0]), rows[pos - 1], nullptr);
  if (st != AOADMM_OK) {
    // copy the message before the handle (its owner) goes away
    char msg[512];
    std::strncpy(msg, aoadmm_last_error(h), sizeof(msg) - 1);
    msg[sizeof(msg) - 1] = 0;
    aoadmm_destroy(h);
    mexErrMsgIdAndTxt(st == AOADMM_ERR_UNSUPPORTED ? "aoadmm:unsupported" : "aoadmm:nvecs", "%s", msg);
  }
  aoadmm_destroy(h);
}
