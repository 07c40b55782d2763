// r1_internal.h -- shared between the translation units of librays1_b200.so; not part of the ABI.
#pragma once

// records the message r1_last_error() returns (per thread) and hands `code` back, so that `return r1_set_error(...)` reads well
int r1_set_error(int code, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
