# usage: build_warp_variant.sh NAME "-DVFI_WARPF_PPT=4 ..."  -> video-frame-interpolation_b200/variants/libvfi_NAME.so
# (warp.cu recompiled with the flags, the other objects taken from the default build)
set -e
cd "$(dirname "$0")/../video-frame-interpolation_b200"
mkdir -p variants csrc/_obj
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -Xptxas -v $2 -c csrc/warp.cu -o csrc/_obj/warp_$1.o 2>&1 | grep -A1 "warp_fwd_fast_kernelI13__nv_bfloat16S2_Lb0ELb0" | grep -o "Used [0-9]* registers" | head -1
nvcc -shared -o variants/libvfi_$1.so csrc/_obj/abi.o csrc/_obj/warp_$1.o csrc/_obj/dcn_simt.o csrc/_obj/dcn_tc.o -gencode arch=compute_100a,code=sm_100a
echo built variants/libvfi_$1.so
