for c in 2 4 8 16; do for co in 0 1; do
  if [ $co = 1 ]; then export FCV_COPY_ONLY=1; else unset FCV_COPY_ONLY; fi
  FCV_CHUNKS=$c python bench.py --no-cpu-baseline --steps 100 | python -c "import json,sys,os; d=json.loads(sys.stdin.read()); print('chunks',os.environ.get('FCV_CHUNKS'),'copyonly',os.environ.get('FCV_COPY_ONLY'),'e2e ms', round(d['e2e']['ms_per_step'],3))"
done; done
