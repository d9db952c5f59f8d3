// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product; never linked into libvqcuda.
//
// postprocess_ref_shim.cpp — C entry points over the REFERENCE'S OWN FPostProcessParameters methods
// (Source/Engine/PostProcess/PostProcess.cpp:37-99 compiled unmodified and in place): the engine-side wrappers that fill the
// EASU / RCAS constant blocks the kernels consume, and the sharpness-stops <-> linear conversion of the UI.
// Output: oracle/_ref/libvqppref.so.
#include "Engine/PostProcess/PostProcess.h"
#include <cstring>

extern "C" {
void vqpp_easu(unsigned con[16], unsigned in_w, unsigned in_h, unsigned cont_w, unsigned cont_h, unsigned out_w, unsigned out_h) {
    FPostProcessParameters::FFSR1_EASU p{};
    p.UpdateEASUConstantBlock(in_w, in_h, cont_w, cont_h, out_w, out_h);
    std::memcpy(con, p.EASUConstantBlock, sizeof(p.EASUConstantBlock));
}
void vqpp_rcas(unsigned con[4], float stops) {
    FPostProcessParameters::FFSR1_RCAS p{};
    p.RCASSharpnessStops = stops;
    p.UpdateRCASConstantBlock();
    std::memcpy(con, p.RCASConstantBlock, sizeof(p.RCASConstantBlock));
}
float vqpp_rcas_linear_from_stops(float stops) { FPostProcessParameters::FFSR1_RCAS p{}; p.RCASSharpnessStops = stops; return p.GetLinearSharpness(); }
float vqpp_rcas_stops_from_linear(float linear) { FPostProcessParameters::FFSR1_RCAS p{}; p.SetLinearSharpness(linear); return p.RCASSharpnessStops; }
}
