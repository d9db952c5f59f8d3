// ORACLE — TEST INFRASTRUCTURE ONLY. Stand-in for the reference's MSVC-only Libs/VQUtils/Include/Log.h (token pasting, sprintf_s)
// so that its Source/Image.cpp compiles with g++ UNMODIFIED and in place: same names, messages go to stderr.
#pragma once
#include <cmath>
#include <cstdio>
#include <string>
#define sqrtf sqrt      /* Image.cpp:68 writes std::sqrtf, which libstdc++ does not declare; std::sqrt(float) is the same function */
namespace Log {
template <class... A> inline void Error(const char* f, A... a)   { std::fprintf(stderr, "[ref Image.cpp] error: ");   std::fprintf(stderr, f, a...); std::fputc('\n', stderr); }
template <class... A> inline void Warning(const char* f, A... a) { std::fprintf(stderr, "[ref Image.cpp] warning: "); std::fprintf(stderr, f, a...); std::fputc('\n', stderr); }
template <class... A> inline void Info(const char* f, A... a)    { (void)f; ((void)a, ...); }
}
