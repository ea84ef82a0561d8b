/* Pins oc_cosf/oc_sinf (orb_oracle.c) to the host libm: exhaustive over every float in [0, 2*pi + eps]
 * (the only inputs computeOrbDescriptor can produce: angle in [0,360] degrees times (float)(pi/180)).
 * Usage: check_sincos [stride]   (stride 1 = exhaustive, ~1.09e9 values). Exit code 0 iff no mismatch. */
#include "orb_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
int main(int argc, char** argv)
{
    unsigned stride = argc > 1 ? (unsigned)atoi(argv[1]) : 1;
    float hi = 6.2832f; uint32_t uhi; memcpy(&uhi, &hi, 4);
    unsigned long long n = 0, mc = 0, ms = 0;
    for (uint64_t u = 0; u <= uhi; u += stride) {
        uint32_t b = (uint32_t)u; float x; memcpy(&x, &b, 4);
        float c1 = cosf(x), s1 = sinf(x), c2 = oc_cosf(x), s2 = oc_sinf(x);
        mc += memcmp(&c1, &c2, 4) != 0; ms += memcmp(&s1, &s2, 4) != 0; n++;
    }
    printf("{\"checked\": %llu, \"cos_mismatch\": %llu, \"sin_mismatch\": %llu}\n", n, mc, ms);
    return (mc || ms) ? 1 : 0;
}
