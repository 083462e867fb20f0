"""Time-step controllers (mirror of src/spitfire/time/stepcontrol.py:14-101): same arithmetic, so adaptive step
sequences reproduce the reference's."""
import numpy as np


class ConstantTimeStep(object):
    """fixed step size (stepcontrol.py:14-48)"""

    def __init__(self, step_size):
        self.step_size = step_size

    def __call__(self, *args, **kwargs):
        return self.step_size

    def first_step_size(self):
        return self.step_size

    def last_step_size(self):
        return self.step_size

    def target_error(self):
        return -1.

    def step_size_is_constant(self):
        return True


class PIController(object):
    """proportional-integral control of the embedded error estimate (stepcontrol.py:51-117)

    dt_new = min(dt * min(max_ramp, (tol/err)^ki * (err_prev/err)^kp), max_step); when the error estimate vanishes the
    step grows by max_ramp."""

    def __init__(self, kp=0.06666666667, ki=0.1333333333, target_error=1.e-4, max_step=1.e4, max_ramp=1.1,
                 first_step=1.e-3):
        self._kp, self._ki = kp, ki
        self._target_error, self._max_step, self._max_ramp, self._first_step = target_error, max_step, max_ramp, \
            first_step
        self._err_history = np.zeros(2)
        self._step_history = np.zeros(2)

    def __call__(self, step_count, step, step_output, *args, **kwargs):
        error = step_output.temporal_error
        if error < 1.e-16:
            return min(step * self._max_ramp, self._max_step)
        if step_count < 1:
            self._err_history[step_count] = error
            self._step_history[step_count] = step
        else:
            self._err_history[0], self._step_history[0] = self._err_history[1], self._step_history[1]
            self._err_history[1], self._step_history[1] = error, step
        ratio = (self._target_error / error) ** self._ki
        if step_count != 0:
            ratio = ratio * (self._err_history[-1] / error) ** self._kp
        return min(step * min(self._max_ramp, ratio), self._max_step)

    def first_step_size(self):
        return self._first_step

    def last_step_size(self):
        return self._step_history[-1]

    def target_error(self):
        return self._target_error

    def step_size_is_constant(self):
        return False
