import cProfile, pstats, sys, os
sys.path.insert(0,'.'); sys.argv=['x','14']
import runpy
cProfile.run("runpy.run_path('tools/profile_hp.py', run_name='__main__')", '/tmp/hp.prof')
p=pstats.Stats('/tmp/hp.prof'); p.sort_stats('cumulative').print_stats(28)
