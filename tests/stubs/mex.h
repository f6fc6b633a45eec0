/* Minimal stand-in for MATLAB's mex.h / matrix.h: declarations only, so that tests can COMPILE (not link or run)
 * matlab-code_b200/matlab/aoadmm_mex.cpp in a container without MATLAB.  Signatures follow the documented
 * C Matrix API (R2018a interleaved-complex-agnostic subset used by the gateway). */
#ifndef STUB_MEX_H_
#define STUB_MEX_H_
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef size_t mwIndex;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
bool mxIsStruct(const mxArray*);
bool mxIsCell(const mxArray*);
bool mxIsChar(const mxArray*);
bool mxIsNumeric(const mxArray*);
bool mxIsDouble(const mxArray*);
bool mxIsComplex(const mxArray*);
bool mxIsSparse(const mxArray*);
bool mxIsEmpty(const mxArray*);
bool mxIsLogical(const mxArray*);
bool mxIsLogicalScalarTrue(const mxArray*);
typedef bool mxLogical;
mxLogical* mxGetLogicals(const mxArray*);
int mexCallMATLAB(int nlhs, mxArray* plhs[], int nrhs, mxArray* prhs[], const char* name);
size_t mxGetNumberOfElements(const mxArray*);
size_t mxGetM(const mxArray*);
mwSize mxGetNumberOfDimensions(const mxArray*);
const mwSize* mxGetDimensions(const mxArray*);
size_t mxGetN(const mxArray*);
void mxSetN(mxArray*, mwSize);
double mxGetScalar(const mxArray*);
double* mxGetPr(const mxArray*);
double mxGetNaN(void);
const char* mxGetClassName(const mxArray*);
mxArray* mxGetField(const mxArray*, mwIndex, const char*);
void mxSetField(mxArray*, mwIndex, const char*, mxArray*);
mxArray* mxGetCell(const mxArray*, mwIndex);
mxArray* mxGetProperty(const mxArray*, mwIndex, const char*);
char* mxArrayToString(const mxArray*);
void mxFree(void*);
mxArray* mxDuplicateArray(const mxArray*);
mxArray* mxCreateDoubleMatrix(mwSize, mwSize, mxComplexity);
mxArray* mxCreateDoubleScalar(double);
mxArray* mxCreateString(const char*);
mxArray* mxCreateStructMatrix(mwSize, mwSize, int, const char**);
void mxDestroyArray(mxArray*);
void mexErrMsgIdAndTxt(const char*, const char*, ...) __attribute__((noreturn));
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);
#ifdef __cplusplus
}
#endif
#endif
