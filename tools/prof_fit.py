"""cProfile of one complete-fit bench step (host hot spots of batch_fit.process_batch)."""
import cProfile
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if __name__ == '__main__':
    os.chdir(ROOT)
    sys.argv = ['bench.py', '--mode', 'fit', '--batch', '1024', '--steps', '1', '--warmup', '0',
                '--no-cpu']
    import bench
    cProfile.runctx('bench.main()', {'bench': bench}, {}, '/tmp/fit.prof')
    p = pstats.Stats('/tmp/fit.prof')
    p.sort_stats('tottime').print_stats(40)
