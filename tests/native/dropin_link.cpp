// Link-level drop-in check: this translation unit declares the solver entry EXACTLY as the
// reference's src/PeakSegFPOPLog.h:15 does (C++ linkage, no extern "C") and calls it the way
// src/interface.cpp:10-15 does.  It must link against libpeaksegdisk_b200.so without any shim.
#include <cstdio>
int PeakSegFPOP_disk(char *, char *, char *);

int main(int argc, char **argv) {
  if (argc != 4) { fprintf(stderr, "usage: %s bedGraph penalty db\n", argv[0]); return 100; }
  char *bedGraph = argv[1];
  char *penalty = argv[2];
  char *db = argv[3];
  int status = PeakSegFPOP_disk(bedGraph, penalty, db);
  printf("status=%d\n", status);
  return status;
}
