# policy sweep of the multi-input fast path: BMPC_TPM_CAPS = first-launch refinements, straggler-round refinements, first straggler ADMM chunk
for caps in ${@:-8,6,25 12,8,100}; do
  echo -n "caps $caps: "; BMPC_TPM_CAPS=$caps timeout 200 python tools/bench_configs.py mimo --steps 10 --warmup 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:(round(v,2) if isinstance(v,float) else v) for k,v in d.items() if k in ('ms_per_step','ms_admm_per_step','ms_polish_per_step','mean_rounds','admm_iters_per_solve','polish_steps_per_solve','unsolved')})"
done
