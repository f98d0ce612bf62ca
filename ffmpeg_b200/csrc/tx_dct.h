// tx_dct.h — AV_TX_FLOAT_DCT (tx_dct.cu): DCT-II forward / DCT-III inverse around the real DFT of tx.cu
#pragma once
#include "common.h"

struct TxDct;
// len as given to av_tx_init (the inverse works on 2 * len points, libavutil/tx_template.c:1844-1848); nullptr on failure
TxDct *tx_dct_create(B200Device *dev, int inv, int len, float scale);
void   tx_dct_free(TxDct *p);
bool   tx_dct_length_ok(int inv, int len);
int    tx_dct_points(const TxDct *p);            // floats per transform, in and out
int    tx_dct_launch(TxDct *p, cudaStream_t st, void *out, const void *in, int64_t count, ptrdiff_t out_step, ptrdiff_t in_step);
