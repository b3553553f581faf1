// TEST INFRASTRUCTURE: the few declarations of R's <R.h> that the reference's src/interface.cpp and
// solver sources use, so the UNMODIFIED interface.cpp can be compiled outside R and linked against
// libpeaksegdisk_b200.so (oracle/Makefile, target `interface`).  Implemented by interface_driver.cpp.
#pragma once
#include <cstdio>
#include <cstring>
#include <cmath>
#include <string>
#define Rprintf printf
extern "C" void Rf_error(const char *fmt, ...);
