for m in 1 0 u1; do
  echo "FCB200_CHAIN_PACK=$m"
  FCB200_CHAIN_PACK=$m python bench.py --steps 20 --warmup 3 --no-ops --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value']/1e9, d['roofline']['frac'], d['clocks'])"
done
