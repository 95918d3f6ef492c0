/* MOCK of R.h for `gcc -fsyntax-only` in an image without R (tests/test_abi.py); never shipped to R. */
#ifndef MOCK_R_H
#define MOCK_R_H
#include <stddef.h>
char *R_alloc(size_t n, int size);
#endif
