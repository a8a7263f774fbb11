"""cProfile of tools/bench_worker_e2e.py's second (warm) file with num_data_workers=0: where the host
time of `embedding_worker` goes."""
import cProfile, pstats, runpy, sys
sys.argv = ['bench_worker_e2e.py', *sys.argv[1:]]
pr = cProfile.Profile()
pr.enable()
try:
    runpy.run_path(str(__import__('pathlib').Path(__file__).with_name('bench_worker_e2e.py')), run_name='__main__')
finally:
    pr.disable()
    st = pstats.Stats(pr)
    st.sort_stats('cumulative').print_stats(45)
