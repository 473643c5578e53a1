for i in 1 2 3 4 5 6; do
  D4_E2E_TRACE=1 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-graph > /tmp/o.txt 2> /tmp/e.txt
  ms=$(tail -1 /tmp/o.txt | python -c "import json,sys; print(round(json.loads(sys.stdin.read())['e2e']['ms_per_step'],2))")
  echo "run $i e2e $ms"; grep allocator /tmp/e.txt | tr '\n' ';'; echo
done
