// psd_fast.cuh — interface of the K3 fast path (psd_fast.cu) used by psd.cu
#pragma once
#include "common.cuh"
int psd_fast_supported(int nfft);
int psd_fast_pos_to_freq(int nfft, int p);
size_t psd_fast_table_elems(int nfft);
void psd_fast_fill_table(int nfft, float2 *t);
int psd_fast_launch(int nfft, const void *d_x, int is_complex, const float *d_win, int chunk, int hop, int navg, i64 n_lines,
                    const float2 *d_tw, float *d_part, cudaStream_t st);
