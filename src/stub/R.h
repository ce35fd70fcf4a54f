/* stand-in for R.h (see Rinternals.h in this directory) */
#include "Rinternals.h"
