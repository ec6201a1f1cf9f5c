for o in "patch_request_cost=512" "patch_request_cost=1024" "patch_request_cost=2048" "patch_request_cost=256" "patch_max_owned=480" "patch_max_owned=448" "patch_stages=2"; do
  echo "== $o"; python bench.py --quick --steps 30 --warmup 5 --no-cpu-baseline --opt $o 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['kernel_ms'], d['clocks'])"
done
