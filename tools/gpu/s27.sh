timeout 300 python tools/dense_ab.py 30 prefetch_ahead=0 prefetch_ahead=1 prefetch_ahead=2 prefetch_ahead=0 prefetch_ahead=1 2>&1 | cut -c1-260
