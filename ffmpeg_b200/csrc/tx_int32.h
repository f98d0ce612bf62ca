// tx_int32.h — AV_TX_INT32_FFT / AV_TX_INT32_MDCT (tx_int32.cu), power-of-two lengths
#pragma once
#include "common.h"

struct TxI32;
TxI32 *tx_i32_create(B200Device *dev, int type, int inv, int len, float scale);      // nullptr on failure
void   tx_i32_free(TxI32 *p);
bool   tx_i32_length_ok(int type, int len);
int    tx_i32_launch(TxI32 *p, cudaStream_t st, void *out, const void *in, ptrdiff_t stride, int64_t count, ptrdiff_t out_step, ptrdiff_t in_step);
void   tx_i32_host_fn(TxI32 *p, void *out, void *in, ptrdiff_t stride);             // av_tx_fn shape: HOST pointers, one transform
