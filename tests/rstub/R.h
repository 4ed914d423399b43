/* tests/rstub/R.h — TEST INFRASTRUCTURE ONLY (see Rinternals.h in this directory) */
#ifndef RSTUB_R_H
#define RSTUB_R_H
#include <stdlib.h>
#include <string.h>
#endif
