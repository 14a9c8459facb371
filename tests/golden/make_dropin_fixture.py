"""TEST INFRASTRUCTURE -- writes tests/golden/dropin_loop.txt: the body of the optimisation loop of the reference's
scripts/model_poses_learning (lines 119-135: corrected poses -> global cloud -> features -> loss -> backward -> step),
VERBATIM, as a test vector.  tests/test_dropin.py executes these very lines with `depth_correction` aliased to
`depth_correction_b200` on the GPU box, where /root/reference does not exist;
tests/test_oracle.py::test_dropin_fixture_is_the_reference_loop checks here (CPU container) that the fixture still is
the reference's text.

    python tests/golden/make_dropin_fixture.py
"""
import os
import textwrap

SCRIPT = '/root/reference/scripts/model_poses_learning'
FIRST, LAST = 'train_poses_corr = create_corrected_poses(train_poses, train_pose_deltas, cfg)', 'optimizer.step()'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'dropin_loop.txt')


def loop_body():
    lines = open(SCRIPT).read().split('\n')
    a = next(i for i, l in enumerate(lines) if l.strip() == FIRST)
    b = next(i for i in range(a, len(lines)) if lines[i].strip() == LAST)
    return textwrap.dedent('\n'.join(lines[a:b + 1])) + '\n', a + 1, b + 1


if __name__ == '__main__':
    body, a, b = loop_body()
    open(OUT, 'w').write(body)
    print('wrote %s (scripts/model_poses_learning:%d-%d)' % (OUT, a, b))
