/* TEST INFRASTRUCTURE ONLY.  Declarations standing in for Biostrings' C interface so that the reference's
 * src/DNA_input.cpp compiles; the oracle only ever feeds character vectors (string_input), so the S4 path these
 * belong to is never taken and the definitions in oracle/umi_ref_driver.cpp just raise. */
#ifndef SARLACC_ORACLE_RSHIM_BIOSTRINGS_H
#define SARLACC_ORACLE_RSHIM_BIOSTRINGS_H

typedef struct { const char* ptr; int length; } Chars_holder;
typedef struct { int length; const void* opaque; } XStringSet_holder;

XStringSet_holder hold_XStringSet(struct RshimObject* x);
int get_length_from_XStringSet_holder(const XStringSet_holder* x);
Chars_holder get_elt_from_XStringSet_holder(const XStringSet_holder* x, int i);
char DNAdecode(char code);

#endif
