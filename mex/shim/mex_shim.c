/* Test-only implementation of the handful of mx* / mex* functions declared in shim/mex.h, backed by
 * malloc.  mexErrMsgIdAndTxt records the message and longjmps back to ofdm_mex_shim_call(), which
 * is how the Python tests drive mexFunction through ctypes on the GPU box. */
#include "mex.h"
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct mxArray_tag {
    mwSize m, n;
    int is_complex, is_char;
    double* re;               /* interleaved (re,im) pairs when complex and !split */
    double* im;               /* split mode only */
    char* str;
};
static jmp_buf g_jmp;
static int g_jmp_armed = 0;
static char g_err[1024];
static void (*g_atexit)(void) = NULL;

mxArray* mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c) {
    mxArray* a = (mxArray*)calloc(1, sizeof(mxArray));
    a->m = m; a->n = n; a->is_complex = (c == mxCOMPLEX);
    size_t cnt = (m != 0 && n != 0) ? m * n : 1;
#ifdef OFDM_MEX_SPLIT_COMPLEX
    a->re = (double*)calloc(cnt, sizeof(double));
    if (a->is_complex) a->im = (double*)calloc(cnt, sizeof(double));
#else
    a->re = (double*)calloc(cnt * (a->is_complex ? 2 : 1), sizeof(double));
#endif
    return a;
}
mxArray* mxCreateDoubleScalar(double v) { mxArray* a = mxCreateDoubleMatrix(1, 1, mxREAL); a->re[0] = v; return a; }
mxArray* mxCreateString(const char* s) {
    mxArray* a = (mxArray*)calloc(1, sizeof(mxArray));
    a->is_char = 1; a->m = 1; a->n = strlen(s); a->str = strdup(s);
    return a;
}
void mxDestroyArray(mxArray* a) { if (!a) return; free(a->re); free(a->im); free(a->str); free(a); }
mwSize mxGetM(const mxArray* a) { return a->m; }
mwSize mxGetN(const mxArray* a) { return a->n; }
size_t mxGetNumberOfElements(const mxArray* a) { return a->m * a->n; }
int mxIsComplex(const mxArray* a) { return a->is_complex; }
int mxIsChar(const mxArray* a) { return a->is_char; }
int mxIsDouble(const mxArray* a) { return !a->is_char; }
double mxGetScalar(const mxArray* a) { return a->re ? a->re[0] : 0.0; }
int mxGetString(const mxArray* a, char* buf, mwSize buflen) {
    if (!a->is_char || strlen(a->str) + 1 > buflen) return 1;
    strcpy(buf, a->str);
    return 0;
}
double* mxGetPr(const mxArray* a) { return a->re; }
#ifdef OFDM_MEX_SPLIT_COMPLEX
double* mxGetPi(const mxArray* a) { return a->im; }
#else
double* mxGetDoubles(const mxArray* a) { return a->re; }
mxComplexDouble* mxGetComplexDoubles(const mxArray* a) { return (mxComplexDouble*)a->re; }
#endif
void mexErrMsgIdAndTxt(const char* id, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    int n = snprintf(g_err, sizeof g_err, "%s: ", id);
    vsnprintf(g_err + n, sizeof g_err - n, fmt, ap);
    va_end(ap);
    if (g_jmp_armed) longjmp(g_jmp, 1);
    fprintf(stderr, "%s\n", g_err);
    abort();
}
int mexAtExit(void (*fn)(void)) { g_atexit = fn; return 0; }

/* ---- test driver entry points (ctypes) ---- */
int ofdm_mex_shim_call(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    g_err[0] = 0;
    g_jmp_armed = 1;
    if (setjmp(g_jmp)) { g_jmp_armed = 0; return 1; }
    mexFunction(nlhs, plhs, nrhs, prhs);
    g_jmp_armed = 0;
    return 0;
}
const char* ofdm_mex_shim_error(void) { return g_err; }
void ofdm_mex_shim_exit(void) { if (g_atexit) g_atexit(); g_atexit = NULL; }
void ofdm_mex_shim_set(mxArray* a, size_t i, double re, double im) {
#ifdef OFDM_MEX_SPLIT_COMPLEX
    a->re[i] = re; if (a->is_complex) a->im[i] = im;
#else
    if (a->is_complex) { a->re[2 * i] = re; a->re[2 * i + 1] = im; } else a->re[i] = re;
#endif
}
void ofdm_mex_shim_get(const mxArray* a, size_t i, double* re, double* im) {
#ifdef OFDM_MEX_SPLIT_COMPLEX
    *re = a->re[i]; *im = a->is_complex ? a->im[i] : 0.0;
#else
    if (a->is_complex) { *re = a->re[2 * i]; *im = a->re[2 * i + 1]; } else { *re = a->re[i]; *im = 0.0; }
#endif
}
