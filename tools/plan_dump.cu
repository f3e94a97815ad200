// Offline view of the conv_umma2 planner's choices (no GPU needed: the plan is printed before the tensor maps are encoded).
//   nvcc -o /tmp/plan_dump tools/plan_dump.cu iris_tts_b200/build/*.o -Iiris_tts_b200/csrc && HFG_U2_VERBOSE=1 /tmp/plan_dump
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "hfg_internal.h"

using namespace hfg;

int main() {
    const int B = getenv("PD_B") ? atoi(getenv("PD_B")) : 16;
    struct Case { int C, k, d, L; } cases[] = {{256, 3, 1, 6896}, {256, 11, 5, 6896}, {128, 3, 1, 55168}, {128, 7, 3, 55168}, {128, 11, 5, 55168},
                                               {64, 3, 1, 110336}, {64, 7, 3, 110336}, {64, 11, 5, 110336}, {32, 3, 1, 220672}, {32, 11, 5, 220672}};
    for (int planes = 1; planes <= 2; ++planes)
        for (int res = 0; res <= 1; ++res)
            for (const Case& c : cases) {
                UmmaConvParams p;
                memset(&p, 0, sizeof p);
                p.g.B = B; p.g.Lin = p.g.Lout = p.g.Mrows = c.L; p.g.Cin = p.g.Cout = p.g.Np = c.C;
                p.g.taps = c.k; p.g.tap_step = res ? 1 : c.d; p.g.tap_off0 = -(c.k - 1) / 2 * p.g.tap_step; p.g.ups_s = 1;
                p.cin_pad = c.C; p.kc = c.C >= 64 ? 64 : 32; p.npass = planes == 2 ? 3 : 1;
                p.y_act = (__nv_bfloat16*)0x1000; p.y_act_lo = (__nv_bfloat16*)0x1000;
                if (res) { p.res_hi = (const __nv_bfloat16*)0x1000; p.res_lo = (const __nv_bfloat16*)0x1000; }
                Umma2Launch L;
                plan_conv_umma2(&L, p, (const __nv_bfloat16*)0x1000, (const __nv_bfloat16*)0x1000, (const __nv_bfloat16*)0x1000,
                                (const __nv_bfloat16*)0x1000, 148);
            }
    return 0;
}
