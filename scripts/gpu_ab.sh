# usage: gpu_ab.sh NAME   -- bench A/B of the default library against variants/libvfi_NAME.so (interleaved, two rounds)
mkdir -p gpurun_out
V=$PWD/video-frame-interpolation_b200/variants/libvfi_$1.so
VFI_B200_LIB=$V timeout 300 python -m pytest tests -x -q -m gpu -k "fused or hot_path or tensor_core_path" 2>&1 | tail -2
for r in 1 2; do
for lib in "" "$V"; do
VFI_B200_LIB=$lib timeout 200 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('${lib:-default}'[-16:], d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'])"
done; done
