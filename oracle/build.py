"""Build the oracle's native pieces (checker only).

  oracle/_build/libplant_oracle.so   gcc build of the C restatement oracle/plant/plant_oracle.c, the episode port
                                     (episode.c) and the kernel-order actor (actor_kernel_order.c)
  oracle/_ref/citation_<variant>.so  byte copies of the reference's own plant binaries (only when
                                     /root/reference exists, i.e. in the build container; the GPU box uses the
                                     copies that travelled with the snapshot)
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ENVS = '/root/reference/envs'
VARIANTS = ['h2000_v90', 'ice', 'cg', 'cg_for', 'h2000_v150', 'h10000_v90']
REF_ONLY = ['cg_timed', 'gust', 'test']       # time-triggered builds: checked against their binaries only (no C restatement)
LIB = os.path.join(HERE, '_build', 'libplant_oracle.so')


def build(force=False):
    src = os.path.join(HERE, 'plant', 'plant_oracle.c')
    epi = os.path.join(HERE, 'plant', 'episode.c')
    ako = os.path.join(HERE, 'plant', 'actor_kernel_order.c')
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    deps = [src, epi, ako, os.path.join(HERE, 'plant', 'plant_support.h')] + \
        [os.path.join(HERE, 'plant', 'gen', f) for f in os.listdir(os.path.join(HERE, 'plant', 'gen'))]
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(['gcc', '-O2', '-ffp-contract=off', '-fopenmp', '-fPIC', '-shared', '-o', LIB, src, epi, ako, '-lm'])
    if os.path.isdir(REF_ENVS):
        os.makedirs(os.path.join(HERE, '_ref'), exist_ok=True)
        for v in VARIANTS + REF_ONLY:
            dst = os.path.join(HERE, '_ref', 'citation_%s.so' % v)
            if not os.path.exists(dst):
                shutil.copy(os.path.join(REF_ENVS, v, '_citation.cpython-38-x86_64-linux-gnu.so'), dst)
                os.chmod(dst, 0o755)
    return LIB


def have_ref():
    return all(os.path.exists(os.path.join(HERE, '_ref', 'citation_%s.so' % v)) for v in VARIANTS)


if __name__ == '__main__':
    print(build(force=True), 'reference binaries:', have_ref())
