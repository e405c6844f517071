#!/bin/bash
# BASELINE.json configs[4]: nonlinear 3D flap ~20 M DoFs - 48 x 288 x 64 Q2 cells = 21,661,155 DoFs
# (SURVEY's 48x288x60 has an odd third repetition after two coarsenings, i.e. a 337 k-DoF coarsest
# multigrid level; x64 coarsens down to 3x18x4) - implicit IQN-ILS-style coupling emulated by 10
# sub-iterations per time window (checkpoint / restore every pass, many Newton solves per step),
# 8 B200 of one box. A "step" of the bench is one pass, so --steps 10 times exactly one full
# window. Run under gpurun --gpus 8.
N=${N:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
  --master-port ${PORT:-29533} bench.py --gpus $N --reps 48,288,64 --n-sub 10 \
  --steps ${STEPS:-10} --warmup ${WARMUP:-10} --no-cpu-baseline --no-variants --no-strong
