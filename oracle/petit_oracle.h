/* Plain-C restatement of the reference's dequantise + GEMM oracle.
 * TEST INFRASTRUCTURE ONLY -- see oracle/petit_oracle.py for the citations and
 * the pinning status.  Only tests/, __graft_entry__.smoke() and bench.py's CPU
 * baseline legs may link or call this. */
#ifndef PETIT_ORACLE_H_
#define PETIT_ORACLE_H_
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
float petit_oracle_e2m1(unsigned code);                 /* test_fp4_gemm_quark.py:10-14 */
float petit_oracle_e4m3(uint8_t bits);                  /* lib/tests/floating_points.h:21-75 */
float petit_oracle_e8m0(uint8_t bits);                  /* dequant.cuh:197-203: bf16 (s<<7) */
uint16_t petit_oracle_f32_to_bf16(float f);             /* round to nearest even */
uint16_t petit_oracle_f32_to_f16(float f);
float petit_oracle_bf16_to_f32(uint16_t b);
float petit_oracle_f16_to_f32(uint16_t h);
/* q: [n, k/2] bytes, low nibble = even k; scales [n, k/16] e4m3 / [n, k/32] e8m0;
 * out [n, k] fp32 (test_fp4_gemm_quark.py:9-20; quantization_utils.cu:405-432) */
void petit_oracle_dequant_nvfp4(float *out, const uint8_t *q, const uint8_t *scales, size_t n, size_t k);
void petit_oracle_dequant_mxfp4(float *out, const uint8_t *q, const uint8_t *scales, size_t n, size_t k);
/* c[m,n] = sum_k a[m,k] * w[n,k], fp32 accumulate (test_fp4_gemm_quark.py:23-24);
 * uses OpenMP threads when compiled with -fopenmp */
void petit_oracle_gemm_f32(float *c, const float *a, const float *w, size_t m, size_t n, size_t k);
#ifdef __cplusplus
}
#endif
#endif
