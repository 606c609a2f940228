// Fast path of the M1 RX chain (Nfft = 4096): placeholder until the radix-16 kernel lands.
#include "common.cuh"
int ofdm_rx_chain_fast4096(ofdm_ctx* ctx, const ofdm_link_params* lp, const void* rx, int64_t B, const uint32_t* tx_bits, uint32_t* out_bits,
                           void* H, int64_t* counts, int32_t* err_stream, double near_eps, bool* handled) {
    *handled = false;
    return OFDM_OK;
}
