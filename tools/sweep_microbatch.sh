for w in 20000 5000 2000 1000; do echo "wait_us=$w"; python tools/bench_microbatch.py --max-wait-us $w 2>/dev/null | tail -1 | cut -c1-900; done
