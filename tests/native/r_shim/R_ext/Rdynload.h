// TEST INFRASTRUCTURE: the routine-registration types of R's <R_ext/Rdynload.h> (see ../R.h).
#pragma once
typedef enum { FALSE = 0, TRUE } Rboolean;
typedef void *(*DL_FUNC)();
typedef unsigned int R_NativePrimitiveArgType;
typedef struct {
  const char *name;
  DL_FUNC fun;
  int numArgs;
  R_NativePrimitiveArgType *types;
} R_CMethodDef;
typedef struct _DllInfo DllInfo;
extern "C" int R_registerRoutines(DllInfo *info, const R_CMethodDef *const cRoutines, const void *callRoutines,
                                  const void *fortranRoutines, const void *externalRoutines);
extern "C" Rboolean R_useDynamicSymbols(DllInfo *info, Rboolean value);
