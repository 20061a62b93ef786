import os, ctypes, multiprocessing as mp, tempfile
libc = ctypes.CDLL(None, use_errno=True)
SYS_pidfd_open, SYS_pidfd_getfd = 434, 438
def child(pid, fd, q):
    pfd = libc.syscall(SYS_pidfd_open, pid, 0)
    if pfd < 0:
        q.put(("pidfd_open failed", ctypes.get_errno())); return
    nfd = libc.syscall(SYS_pidfd_getfd, pfd, fd, 0)
    if nfd < 0:
        q.put(("pidfd_getfd failed", ctypes.get_errno())); return
    q.put(("ok", os.pread(nfd, 5, 0)))
if __name__ == "__main__":
    f = tempfile.TemporaryFile(); f.write(b"hello"); f.flush()
    ctx = mp.get_context("spawn"); q = ctx.Queue()
    p = ctx.Process(target=child, args=(os.getpid(), f.fileno(), q)); p.start(); print(q.get(timeout=30)); p.join()
