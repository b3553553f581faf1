// TEST INFRASTRUCTURE: the one SEXP type code src/interface.cpp names (see R.h).
#pragma once
#define STRSXP 16
