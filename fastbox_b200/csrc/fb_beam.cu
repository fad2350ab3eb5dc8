// fb_beam.cu -- per-channel zero-padded 2-D FFT beam convolution (beams.py:81-87).
#include "fb_launch.h"

using namespace fb;

extern "C" int fb_beam_convolve(fb_plan* p, const float* beam, const float* field, float* out) {
    (void)p; (void)beam; (void)field; (void)out;
    set_error("fb_beam_convolve: not implemented yet");
    return -5;
}
