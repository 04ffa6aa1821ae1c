// Host-side weight packing: torch state_dict tensors -> the device layout of rc_linear.cuh.
// Pure C++ (also compiled by tests/host_harness.cpp to check the layout conventions on the CPU).
#pragma once
#include <string.h>
#include <vector>

// Linear [out, in] -> [out_pad, in_pad], zero padded (K padded to a multiple of 16, rows to a multiple of 4).
inline void rc_pack_linear(const float* w, const float* b, int out, int in, int out_pad, int in_pad,
                           std::vector<float>& pw, std::vector<float>& pb) {
    pw.assign((size_t)out_pad * in_pad, 0.f);
    pb.assign((size_t)out_pad, 0.f);
    for (int o = 0; o < out; ++o) {
        memcpy(&pw[(size_t)o * in_pad], w + (size_t)o * in, (size_t)in * sizeof(float));
        pb[o] = b[o];
    }
}

// torch.nn.LSTM layer (weight_ih [4H,H], weight_hh [4H,H], gate order i,f,g,o in blocks of H rows) ->
// gate-interleaved [4H, 2H]: row 4j+g = [weight_ih[g*H+j, :] | weight_hh[g*H+j, :]], bias = b_ih + b_hh.
inline void rc_pack_lstm(const float* wih, const float* whh, const float* bih, const float* bhh, int H,
                         std::vector<float>& pw, std::vector<float>& pb) {
    pw.resize((size_t)4 * H * 2 * H);
    pb.resize((size_t)4 * H);
    for (int j = 0; j < H; ++j)
        for (int g = 0; g < 4; ++g) {
            const size_t src = (size_t)(g * H + j), dst = (size_t)(4 * j + g);
            memcpy(&pw[dst * 2 * H], wih + src * H, (size_t)H * sizeof(float));
            memcpy(&pw[dst * 2 * H + H], whh + src * H, (size_t)H * sizeof(float));
            pb[dst] = bih[src] + bhh[src];
        }
}
