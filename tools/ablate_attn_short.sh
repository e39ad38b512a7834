for D in 0 1 4 8 12 16 32; do echo "== dbg=$D"; B200PF_ATTN_DBG=$D python tools/bench_attn.py 1024 10 2>&1 | tail -4 | python -c "
import json,sys
print(' '.join('%s %.4f' % (json.loads(l)['case'].replace('config','c'), json.loads(l)['ms']) for l in sys.stdin))"; done
