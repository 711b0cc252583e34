/* minimal stand-in for R's headers: ONLY for `gcc -fsyntax-only` of the rglue C sources in an image without R */
#ifndef MOCK_R_H
#define MOCK_R_H
#include <stddef.h>
#endif
