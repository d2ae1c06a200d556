"""Stage the reference's OWN task modules for the shadow-module integration test.

Copies tasks/repeatability.py, tasks/MHA.py and utils/visualization.py -- and deliberately NOT utils/extracter.py,
utils/matcher.py, utils/projection.py, which the test replaces with this package's drop-ins -- from the reference tree
(default /root/reference, or KB_REFERENCE_ROOT) into baseline/_ref/, which is git-ignored (no reference source enters
the history) but travels to the GPU box with the working tree.  tests/test_shadow_integration.py skips when neither
baseline/_ref nor a reference tree is present."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get('KB_REFERENCE_ROOT', '/root/reference')
FILES = ['tasks/repeatability.py', 'tasks/MHA.py', 'utils/visualization.py']


def main():
    if not os.path.isfile(os.path.join(SRC, FILES[0])):
        print(f'no reference tree at {SRC}: nothing staged')
        return 1
    dst = os.path.join(ROOT, 'baseline', '_ref')
    for f in FILES:
        os.makedirs(os.path.dirname(os.path.join(dst, f)), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, f), os.path.join(dst, f))
    print(f'staged {len(FILES)} reference modules under {dst}')
    return 0


if __name__ == '__main__':
    sys.exit(main())
