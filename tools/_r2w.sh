mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for c in syn-paths wd-movies; do WANDB_MODE=disabled timeout 300 python -m kgvae.experiments.train --config configs/autoreg_$c.yaml --max-epochs 2 --checkpoint-dir /tmp/ck_$c 2>&1 | grep -v "^\[log\]" | tail -6; echo "train $c rc=$?"; done
ls -la /tmp/ck_syn-paths/*/ | head
