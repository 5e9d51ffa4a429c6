#!/bin/bash
# Build the csrc tree of a git revision (default HEAD) into srgan_st_b200/libsrst_ab.so so that
# tools/sweep_st.py can A/B it against the working-tree build on the same GPU box (SRST_LIB=...).
REV=${1:-HEAD}
T=$(mktemp -d)
mkdir -p $T/a/b $T/include
git show $REV:include/srst.h > $T/include/srst.h
for f in $(git ls-tree --name-only $REV srgan_st_b200/csrc/); do git show $REV:$f > $T/a/b/$(basename $f); done
(cd $T/a/b && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --shared -Xcompiler -fPIC -o /root/repo/srgan_st_b200/libsrst_ab.so srst_cabi.cu) && echo "built libsrst_ab.so from $REV"
rm -rf $T
