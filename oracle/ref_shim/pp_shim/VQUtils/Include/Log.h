// ORACLE — TEST INFRASTRUCTURE ONLY. Stand-in for the MSVC-only VQUtils Log.h that the reference's
// Source/Engine/PostProcess/PostProcess.cpp reaches as "../../VQUtils/Include/Log.h" (resolved through -I .../pp_shim/a/b).
#pragma once
#include <cmath>
#include <cstdio>
#define powf pow        /* PostProcess.cpp:37-38 writes std::powf / std::log10f, which libstdc++ does not declare; */
#define log10f log10    /* std::pow(float,float) / std::log10(float) are the float overloads                        */
namespace Log {
template <class... A> inline void Info(const char*, A...) {}
template <class... A> inline void Warning(const char*, A...) {}
template <class... A> inline void Error(const char*, A...) {}
}
