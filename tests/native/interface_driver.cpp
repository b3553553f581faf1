// TEST INFRASTRUCTURE.  Plays R for the reference's own src/interface.cpp, compiled UNMODIFIED from
// /root/reference and linked against libpeaksegdisk_b200.so instead of the reference's solver:
//   1. "loads the package": calls R_init_PeakSegDisk(), which registers the .C routines
//      (src/interface.cpp:58-73);
//   2. looks up "PeakSegFPOP_interface" among the registered routines (3 x STRSXP) and calls it the
//      way .C() does from R/PeakSegFPOP_file.R:66-71: three char** vectors of length one;
//   3. Rf_error() prints "Error: <message>" and ends the process with status 1, like R's error.
// usage: interface_driver <bedGraph> <penalty> <db>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "r_shim/R.h"
#include "r_shim/R_ext/Rdynload.h"
#include "r_shim/Rinternals.h"

struct _DllInfo { const R_CMethodDef *c_routines; int dynamic_symbols; };
static _DllInfo g_dll = {nullptr, 1};

extern "C" void Rf_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  fprintf(stderr, "Error: ");
  vfprintf(stderr, fmt, ap);
  fprintf(stderr, "\n");
  va_end(ap);
  exit(1);
}
extern "C" int R_registerRoutines(DllInfo *info, const R_CMethodDef *const c, const void *, const void *, const void *) {
  info->c_routines = c;
  return 1;
}
extern "C" Rboolean R_useDynamicSymbols(DllInfo *info, Rboolean value) {
  const Rboolean old = info->dynamic_symbols ? TRUE : FALSE;
  info->dynamic_symbols = value;
  return old;
}
extern "C" void R_init_PeakSegDisk(DllInfo *info);

int main(int argc, char **argv) {
  if (argc != 4) { fprintf(stderr, "usage: %s bedGraph penalty db\n", argv[0]); return 100; }
  R_init_PeakSegDisk(&g_dll);
  if (g_dll.dynamic_symbols != 0) { fprintf(stderr, "R_useDynamicSymbols(FALSE) was not called\n"); return 101; }
  const R_CMethodDef *m = g_dll.c_routines;
  for (; m && m->name; m++) if (strcmp(m->name, "PeakSegFPOP_interface") == 0) break;
  if (!m || !m->name) { fprintf(stderr, "PeakSegFPOP_interface is not registered\n"); return 102; }
  if (m->numArgs != 3 || m->types[0] != STRSXP || m->types[1] != STRSXP || m->types[2] != STRSXP) {
    fprintf(stderr, "unexpected registration\n"); return 103;
  }
  char *file_vec[1] = {argv[1]}, *pen_vec[1] = {argv[2]}, *temp_vec[1] = {argv[3]};
  typedef void (*dotC3)(char **, char **, char **);
  ((dotC3)m->fun)(file_vec, pen_vec, temp_vec);
  printf("ok\n");
  return 0;
}
