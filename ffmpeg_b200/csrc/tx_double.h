// tx_double.h — AV_TX_DOUBLE_FFT / AV_TX_DOUBLE_MDCT (tx_double.cu), power-of-two lengths
#pragma once
#include "common.h"

struct TxDbl;
TxDbl *tx_dbl_create(B200Device *dev, int type, int inv, int len, double scale);      // nullptr on failure
void   tx_dbl_free(TxDbl *p);
bool   tx_dbl_length_ok(int type, int len);
int    tx_dbl_launch(TxDbl *p, cudaStream_t st, void *out, const void *in, ptrdiff_t stride, int64_t count, ptrdiff_t out_step, ptrdiff_t in_step);
void   tx_dbl_host_fn(TxDbl *p, void *out, void *in, ptrdiff_t stride);               // av_tx_fn shape: HOST pointers, one transform
