// oracle/shim/pre.h -- TEST INFRASTRUCTURE ONLY (force-included when compiling the reference).
// logging.h:12 defines `lprintf()` as a zero-argument macro which GCC rejects when it is called
// with arguments; pre-empt that header with a variadic no-op of the same name.
#ifndef LOGGING_H
#define LOGGING_H
#include <stdio.h>
extern FILE* logF;
#define lprintf(...)
#endif
