# usage: build_variant.sh NAME "-DV6_ROTATE=1 ..."  -> video-frame-interpolation_b200/variants/libvfi_NAME.so
set -e
cd "$(dirname "$0")/../video-frame-interpolation_b200"
mkdir -p variants csrc/_obj
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr $2 -c csrc/dcn_tc.cu -o csrc/_obj/dcn_tc_$1.o
nvcc -shared -o variants/libvfi_$1.so csrc/_obj/abi.o csrc/_obj/warp.o csrc/_obj/dcn_simt.o csrc/_obj/dcn_tc_$1.o -gencode arch=compute_100a,code=sm_100a
echo built variants/libvfi_$1.so
