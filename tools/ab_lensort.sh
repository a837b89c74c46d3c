for v in 0 1; do echo "== NO_LENSORT=$v"; MOIRA_B200_NO_LENSORT=$v timeout 900 python tools/bench_configs.py 2>&1 | grep -E "C3|C5"; done
