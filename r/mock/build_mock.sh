#!/bin/bash
# Builds r/gpb200_r_mock.so: r/shim.c + the functional mock of the R C API, linked against libgpb200.so.
set -e
cd "$(dirname "$0")/../.."
gcc -O1 -g -shared -fPIC -Wall -Wextra -Wno-cast-function-type -Werror -Ir/mock -Iinclude r/shim.c r/mock/mock_r.c \
    -Lgp_b200/lib -lgpb200 -Wl,-rpath,'$ORIGIN/../gp_b200/lib' -o r/gpb200_r_mock.so
echo r/gpb200_r_mock.so
