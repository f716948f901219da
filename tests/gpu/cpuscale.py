import multiprocessing as mp, time
def work(n):
    s=0
    for i in range(n): s+=i*i%7
    return s
if __name__=="__main__":
    for p in (1,4,8,12,16):
        t=time.time()
        with mp.Pool(p) as pool: pool.map(work,[3_000_000]*p)
        dt=time.time()-t
        print(p,"procs: %.2fs  throughput %.2f units/s"%(dt,p/dt))
