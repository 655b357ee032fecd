#!/bin/bash
# GPU side of the round's profile evidence (run under gpurun from the repo root, after `python bench.py` has exited 0).
#   1. launch list of the bench command (per-launch times are cold-cache and serialised: only shares are meaningful)
#   2. ncu --set full of the dominant kernel (gemm_tn), the catalogue kernels, K1 at catalogue scale, attention, wgrad
# Reports land in gpurun_out/; tools/profile_round.py turns them into profiles/*.txt and profiles/traffic.json.
set -u
TAG=${1:-r2}
PARTS=${2:-"launches gemm topk k1 c3k misc"}          # which captures to take
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-extras"
has() { case " $PARTS " in *" $1 "*) return 0;; *) return 1;; esac; }
has launches && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${TAG}_launches.csv \
  $B --no-catalogue > gpurun_out/${TAG}_ncu_launches.log 2>&1
has gemm && timeout 400 ncu --set full --clock-control none -k regex:gemm_tn -s 22 -c 22 -o gpurun_out/${TAG}_gemm_tn \
  $B --no-catalogue > gpurun_out/${TAG}_ncu_gemm.log 2>&1
has topk && timeout 400 ncu --set full --clock-control none -k regex:catalogue_ -s 2 -c 2 -o gpurun_out/${TAG}_topk \
  $B > gpurun_out/${TAG}_ncu_topk.log 2>&1
has k1 && KB_B=81920 KB_N=1000000 KB_VALID=1.0 timeout 300 ncu --set full --clock-control none -k regex:embed_ln -s 3 -c 1 \
  -o gpurun_out/${TAG}_k1 python tools/kbench.py k1 4 > gpurun_out/${TAG}_ncu_k1.log 2>&1
has c3k && timeout 300 ncu --set full --clock-control none -k "regex:score_kernel|embed_bwd_kernel|adam_kernel" -c 3 \
  -o gpurun_out/${TAG}_c3k env SRFRD_SCALE_ITERS=1 python tools/kbench.py c3scale 1 > gpurun_out/${TAG}_ncu_c3k.log 2>&1
has misc && timeout 300 ncu --set full --clock-control none -k "regex:attn_|wgrad|ln_bwd|ln_fwd|score_kernel|pack_|embed_|dp_adam|adam_fused" -s 40 -c 14 \
  -o gpurun_out/${TAG}_misc $B --no-catalogue > gpurun_out/${TAG}_ncu_misc.log 2>&1
# the reports together exceed what gpurun copies back (64 MiB): summarise them here and keep only the text
python tools/profile_round.py ${TAG} gpurun_out/profiles_${TAG} > gpurun_out/${TAG}_summary.log 2>&1
rm -f gpurun_out/${TAG}_*.ncu-rep
ls -la gpurun_out/ gpurun_out/profiles_${TAG}
