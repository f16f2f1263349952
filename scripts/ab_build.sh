#!/bin/bash
# Builds variants of libptb200.so for same-box A/B runs: scripts/ab_build.sh name "-DPTB_X=0 ..." -> glsl-pathtracer_b200/ab/libptb200_<name>.so
# (select at run time with PTB200_LIB=<path>).  The variants are measurement aids; the shipped library is the default build.
set -e
cd "$(dirname "$0")/../glsl-pathtracer_b200/csrc"
name=$1; shift
mkdir -p ../ab
make -s OUT=../ab LIBNAME=libptb200_$name.so BUILD=build_ab_$name EXTRA="$*"
echo ../ab/libptb200_$name.so
