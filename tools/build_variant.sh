#!/bin/bash
# A/B build of the library with extra nvcc defines: tools/build_variant.sh NAME "-DQZ_SC_TRACE ..." [sources to rebuild]
# -> tools/_libs/libquill_NAME.so (git-ignored; select it with QZ_LIB_PATH).  The other objects are the in-tree ones.
set -e
NAME=$1; DEFS=$2; shift 2
SRCS=${@:-sumcheck.cu}
cd "$(dirname "$0")/../quill_zkvm_b200/csrc"
make -j4 >/dev/null
NCCL_INC=$(python -c "import nvidia.nccl, os; print(os.path.join(list(nvidia.nccl.__path__)[0], 'include'))" 2>/dev/null)
mkdir -p ../../tools/_libs/obj_$NAME
OBJS=""
for f in api.cu sumcheck.cu msm.cu mlpcs.cu comm.cu; do
  if [[ " $SRCS " == *" $f "* ]]; then
    nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr $DEFS \
      ${NCCL_INC:+-I$NCCL_INC} -c -o ../../tools/_libs/obj_$NAME/${f%.cu}.o $f
    OBJS="$OBJS ../../tools/_libs/obj_$NAME/${f%.cu}.o"
  else
    OBJS="$OBJS ${f%.cu}.o"
  fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../tools/_libs/libquill_$NAME.so $OBJS -lcudart -ldl
echo built tools/_libs/libquill_$NAME.so
