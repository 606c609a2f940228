/* Minimal stand-in for MATLAB's / Octave's mex.h, just enough to compile (and, through
 * mex_shim.c, to exercise) mex/ofdm_mex.c in an image that has neither MATLAB nor Octave.
 * It models the R2018a interleaved-complex API by default; define OFDM_MEX_SPLIT_COMPLEX to model
 * the legacy / Octave split-complex API (mxGetPr / mxGetPi).  NOT a replacement for the real header:
 * build against MATLAB's `mex` or Octave's `mkoctfile --mex` in production. */
#ifndef OFDM_SHIM_MEX_H
#define OFDM_SHIM_MEX_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
#ifndef OFDM_MEX_SPLIT_COMPLEX
#define MX_HAS_INTERLEAVED_COMPLEX 1   /* what MATLAB's mex.h defines under `mex -R2018a` */
#endif
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
typedef struct { double real, imag; } mxComplexDouble;

mxArray* mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c);
mxArray* mxCreateDoubleScalar(double v);
mxArray* mxCreateString(const char* s);
void mxDestroyArray(mxArray* a);
mwSize mxGetM(const mxArray* a);
mwSize mxGetN(const mxArray* a);
size_t mxGetNumberOfElements(const mxArray* a);
int mxIsComplex(const mxArray* a);
int mxIsChar(const mxArray* a);
int mxIsDouble(const mxArray* a);
double mxGetScalar(const mxArray* a);
int mxGetString(const mxArray* a, char* buf, mwSize buflen);
double* mxGetPr(const mxArray* a);
#ifdef OFDM_MEX_SPLIT_COMPLEX
double* mxGetPi(const mxArray* a);
#else
double* mxGetDoubles(const mxArray* a);
mxComplexDouble* mxGetComplexDoubles(const mxArray* a);
#endif
void mexErrMsgIdAndTxt(const char* id, const char* fmt, ...);
int mexAtExit(void (*fn)(void));
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);
#ifdef __cplusplus
}
#endif
#endif
