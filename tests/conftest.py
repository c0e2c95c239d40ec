import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run with -m gpu on the B200 box)')


H2O = 'O 0 0 0; H 0 -0.757 0.587; H 0 0.757 0.587'
BENZENE = '''c   1.217739890298750 -0.703062453466927  0.000000000000000
h   2.172991468538160 -1.254577209307266  0.000000000000000
c   1.217739890298750  0.703062453466927  0.000000000000000
h   2.172991468538160  1.254577209307266  0.000000000000000
c   0.000000000000000  1.406124906933854  0.000000000000000
h   0.000000000000000  2.509154418614532  0.000000000000000
c  -1.217739890298750  0.703062453466927  0.000000000000000
h  -2.172991468538160  1.254577209307266  0.000000000000000
c  -1.217739890298750 -0.703062453466927  0.000000000000000
h  -2.172991468538160 -1.254577209307266  0.000000000000000
c   0.000000000000000 -1.406124906933854  0.000000000000000
h   0.000000000000000 -2.509154418614532  0.000000000000000'''

EMU_LIB = os.path.join(ROOT, 'tests', 'emu', 'libb200jk_emu.so')


@pytest.fixture(scope='session')
def emu_lib():
    if not os.path.exists(EMU_LIB):
        import subprocess
        subprocess.check_call(['make', '-s', '-C', os.path.join(ROOT, 'pyscf_b200', 'csrc'), 'emu'])
    return EMU_LIB
