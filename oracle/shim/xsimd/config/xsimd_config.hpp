// ORACLE - TEST INFRASTRUCTURE ONLY.  Stand-in for the configuration header of xsimd 14.3.0
// (third-party, not vendored in the reference tree; CMakeLists.txt:72 pins the version).
// See xsimd.hpp next to this file.
#pragma once
#define XSIMD_VERSION_MAJOR 14
#define XSIMD_VERSION_MINOR 3
#define XSIMD_VERSION_PATCH 0
