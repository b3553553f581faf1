// Stand-in for R's <R.h> so the reference solver's two source files compile outside R.
// The reference uses R.h only for Rprintf (src/PeakSegFPOPLog.cpp:7, src/funPieceListLog.cpp:7).
// TEST INFRASTRUCTURE ONLY (oracle build); never part of the product library.
#pragma once
#include <cstdio>
#include <cstring>
#include <cmath>
#include <string>
#define Rprintf printf
