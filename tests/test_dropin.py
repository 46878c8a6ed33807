"""Drop-in boundary: the reference's OWN base/train.py, unmodified, on top of this repo's modules.

  * CPU, build container (needs /root/reference): train.py is executed by runpy after serl_b200.dropin.install(); every
    import resolves, its argparse namespace is accepted by Parameters, select_env / env.seed / the seeding block run, and it
    stops exactly where the GPU is needed (Agent() -> NativeError "needs a CUDA device"): the product has no CPU fallback.
  * GPU + reference tree present (nowhere in this project's pipeline: the GPU box has no /root/reference): the full run.
  * GPU box: the copy-free launcher examples/train.py drives the same call sequence (base/train.py:54-139) through the same
    classes and must leave the three checkpoint files with the reference's key names."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TRAIN = '/root/reference/base/train.py'
RUN = ("import sys; sys.path.insert(0, %r); import serl_b200.dropin as d; d.install(); import runpy; "
       "sys.argv = ['train.py'] + %r; runpy.run_path(%r, run_name='__main__')")
ARGS = ['-frames', '6000', '-pop_size', '6', '-mut_type', 'normal', '-test_ea']


def check_checkpoints(folder):
    pop = torch.load(os.path.join(folder, 'evo_nets.pkl'), weights_only=False)
    assert sorted(pop) == ['actor_%d' % i for i in range(6)]
    keys = ['net.0.weight', 'net.0.bias', 'net.2.weight', 'net.2.bias', 'net.3.gamma', 'net.3.beta', 'net.5.weight', 'net.5.bias',
            'net.6.gamma', 'net.6.beta', 'net.8.weight', 'net.8.bias', 'net.9.gamma', 'net.9.beta', 'net.11.weight', 'net.11.bias']
    assert list(pop['actor_0']) == keys                                   # base/core/genetic_agent.py:78-101 module tree
    assert list(torch.load(os.path.join(folder, 'elite_net.pkl'), weights_only=False)) == keys
    assert list(torch.load(os.path.join(folder, 'rl_net.pkl'), weights_only=False)) == keys


@pytest.mark.skipif(not os.path.exists(TRAIN), reason='needs the reference tree (build container only)')
@pytest.mark.skipif(torch.cuda.is_available(), reason='CPU leg')
def test_unmodified_reference_train_py_reaches_the_gpu_boundary(tmp_path):
    p = subprocess.run([sys.executable, '-c', RUN % (ROOT, ARGS, TRAIN)], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert p.returncode != 0
    assert 'needs a CUDA device' in p.stderr and 'NativeError' in p.stderr, p.stderr[-2000:]
    assert 'agent.Agent(parameters, env)' in p.stderr                      # base/train.py:95 is where it stopped
    assert os.path.isdir(tmp_path / 'tmp')                                  # Parameters(cla) ran (parameters.py:128-129)


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(TRAIN) or not torch.cuda.is_available(), reason='needs the reference tree next to a GPU')
def test_unmodified_reference_train_py_runs_on_the_gpu(tmp_path):
    p = subprocess.run([sys.executable, '-c', RUN % (ROOT, ARGS, TRAIN)], cwd=tmp_path, capture_output=True, text=True, timeout=1800)
    assert p.returncode == 0, p.stderr[-3000:]
    check_checkpoints(tmp_path / 'tmp')


@pytest.mark.gpu
def test_copy_free_launcher_runs_the_train_py_call_sequence(tmp_path):
    p = subprocess.run([sys.executable, os.path.join(ROOT, 'examples', 'train.py')] + ARGS, cwd=tmp_path, capture_output=True, text=True,
                       timeout=1800)
    assert p.returncode == 0, p.stderr[-3000:]
    assert 'Frames:' in p.stdout
    check_checkpoints(tmp_path / 'tmp')


@pytest.mark.gpu
def test_launcher_with_the_reference_default_proximal_mutation(tmp_path):
    """base/train.py:32 defaults to -mut_type proximal: the Jacobian-scaled mutation runs batched on the device, fed by the
    per-actor device replay buffers the rollout kernel fills."""
    args = ['-frames', '9000', '-pop_size', '6', '-mut_type', 'proximal', '-test_ea']
    p = subprocess.run([sys.executable, os.path.join(ROOT, 'examples', 'train.py')] + args, cwd=tmp_path, capture_output=True, text=True,
                       timeout=1800)
    assert p.returncode == 0, p.stderr[-3000:]
    check_checkpoints(tmp_path / 'tmp')


@pytest.mark.gpu
def test_launcher_with_distillation_crossover_and_safe_mutation(tmp_path):
    """the reference's shipped operator set (SERL50 config.yaml: mut_type safe, distil_type fitness): Q-filtered distillation
    crossover and the safe mutation, both batched on the device."""
    args = ['-frames', '9000', '-pop_size', '8', '-mut_type', 'safe', '-use_distil', '-distil_type', 'fitness', '-test_ea']
    p = subprocess.run([sys.executable, os.path.join(ROOT, 'examples', 'train.py')] + args, cwd=tmp_path, capture_output=True, text=True,
                       timeout=1800)
    assert p.returncode == 0, p.stderr[-3000:]
    pop = torch.load(os.path.join(tmp_path, 'tmp', 'evo_nets.pkl'), weights_only=False)
    assert sorted(pop) == ['actor_%d' % i for i in range(8)]
    assert all(torch.isfinite(v).all() for v in pop['actor_3'].values())


def test_proximal_default_of_train_py_is_kept():
    """base/train.py's CLI default is -mut_type proximal: implemented (serl_b200/evo_prox.py), so Parameters keeps it."""
    import types
    from serl_b200.parameters import Parameters
    cwd = os.getcwd()
    os.chdir('/tmp')
    try:
        a = Parameters(types.SimpleNamespace(mut_type='proximal', pop_size=4))
    finally:
        os.chdir(cwd)
    assert a.mut_type == 'proximal'
