/* stand-in for R_ext/Rdynload.h (see ../Rinternals.h) */
#ifndef GWSEM_STUB_RDYNLOAD_H_
#define GWSEM_STUB_RDYNLOAD_H_
#include "../Rinternals.h"
typedef void* (*DL_FUNC)(void);
typedef struct { const char* name; DL_FUNC fun; int numArgs; } R_CallMethodDef;
typedef struct { const R_CallMethodDef* call; int dynamic; } DllInfo;
int R_registerRoutines(DllInfo* info, const void* c, const R_CallMethodDef* call, const void* f, const void* e);
Rboolean R_useDynamicSymbols(DllInfo* info, Rboolean value);
#endif
