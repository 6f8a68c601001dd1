import sys; sys.path.insert(0,'.')
import quill_zkvm_b200 as q
c=q.Context(0)
print("imad.wide MAC/s %.3e  imad32/s %.3e  fr mul/s %.3e fq mul/s %.3e" % (c.bench_imad(0), c.bench_imad(1), c.bench_fp_mul(0), c.bench_fp_mul(1)))
