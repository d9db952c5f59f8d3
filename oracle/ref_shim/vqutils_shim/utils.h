// ORACLE — TEST INFRASTRUCTURE ONLY. Stand-in for the three DirectoryUtil helpers Libs/VQUtils/Source/Image.cpp calls
// (the reference's utils.cpp is Win32 code).
#pragma once
#include <string>
namespace DirectoryUtil {
inline std::string GetFileExtension(const std::string& p) { const size_t d = p.find_last_of('.'); return d == std::string::npos ? "" : p.substr(d + 1); }
inline std::string GetFolderPath(const std::string& p) { const size_t s = p.find_last_of("/\\"); return s == std::string::npos ? "" : p.substr(0, s + 1); }
inline bool CreateFolderIfItDoesntExist(const std::string&) { return true; }   // tests write into an existing temp directory
}
