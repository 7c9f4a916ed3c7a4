/* TEST INFRASTRUCTURE ONLY -- the slice of Biostrings' C interface r_glue.cpp uses (hold_XStringSet and friends, as in
 * /root/reference/src/DNA_input.cpp:64-75), over the toy objects of tests/rstub/rstub.cpp. */
#ifndef SARLACC_RSTUB_BIOSTRINGS_INTERFACE_H
#define SARLACC_RSTUB_BIOSTRINGS_INTERFACE_H

#include "Rinternals.h"

typedef struct { const char* ptr; int length; } Chars_holder;
typedef struct { int length; const void* opaque; } XStringSet_holder;

XStringSet_holder hold_XStringSet(SEXP x);
int get_length_from_XStringSet_holder(const XStringSet_holder* x);
Chars_holder get_elt_from_XStringSet_holder(const XStringSet_holder* x, int i);
char DNAdecode(char code);

#endif
