// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// ffx_ref_shim.cpp — builds the REFERENCE's own FidelityFX constant-setup functions, from the
// reference's own headers where they lie (-I /root/reference/Shaders), exactly the way the engine
// does (Source/Engine/PostProcess/PostProcess.cpp:21-35: `#define A_CPU 1` then include
// ffx_a.h / ffx_cas.h / ffx_fsr1.h). No reference source is copied into this repo: this file only
// #includes them. Output: oracle/_ref/libffxref.so (git-ignored). Used by tests to pin the
// oracle's restated CasSetup / FsrEasuCon / FsrRcasCon / SpdSetup bit-for-bit.
#include <cstdint>
#include <cstdlib>
#include <cmath>
#define A_CPU 1
#include "AMDFidelityFX/FSR1.0/ffx_a.h"
#include "AMDFidelityFX/CAS/ffx_cas.h"
#include "AMDFidelityFX/FSR1.0/ffx_fsr1.h"
#include "AMDFidelityFX/SPD/ffx_spd.h"

extern "C" {
void ref_cas_setup(uint32_t con[8], float sharpness, float in_w, float in_h, float out_w, float out_h) {
    CasSetup(con, con + 4, sharpness, in_w, in_h, out_w, out_h);
}
void ref_fsr_easu_con(uint32_t con[16], float vp_w, float vp_h, float in_w, float in_h, float out_w, float out_h) {
    FsrEasuCon(con, con + 4, con + 8, con + 12, vp_w, vp_h, in_w, in_h, out_w, out_h);
}
void ref_fsr_rcas_con(uint32_t con[4], float stops) { FsrRcasCon(con, stops); }
void ref_spd_setup(uint32_t dispatch_xy[2], uint32_t wg_offset[2], uint32_t nwg_mips[2], const uint32_t rect[4], int mips) {
    uint32_t r[4] = {rect[0], rect[1], rect[2], rect[3]};
    SpdSetup(dispatch_xy, wg_offset, nwg_mips, r, mips);
}
}
