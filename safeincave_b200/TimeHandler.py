"""Time controllers with the interface of the reference's safeincave/TimeHandler.py
(TimeControllerBase :24-126, TimeController :130-166, TimeControllerParabolic :170-273).
Pure host bookkeeping; times are stored in seconds."""
from abc import ABC, abstractmethod

import numpy as np

from .Utils import day, hour, minute, year

_UNITS = {"second": 1, "minute": minute, "hour": hour, "day": day, "year": year}


class TimeControllerBase(ABC):
    def __init__(self, initial_time, final_time, time_unit="second"):
        if time_unit not in _UNITS:
            raise Exception(f"Time unit {time_unit} not supported.")
        self.time_unit = time_unit
        self.time_conversion = _UNITS[time_unit]
        self.t_final = final_time * self.time_conversion
        self.t_initial = initial_time * self.time_conversion
        self.t = initial_time * self.time_conversion
        self.step_counter = 0

    def keep_looping(self):
        return self.t < self.t_final

    @abstractmethod
    def advance_time(self):
        pass


class TimeController(TimeControllerBase):
    def __init__(self, dt, initial_time, final_time, time_unit="second"):
        super().__init__(initial_time, final_time, time_unit)
        self.dt = dt * self.time_conversion

    def advance_time(self):
        self.step_counter += 1
        self.t += self.dt


class TimeControllerParabolic(TimeControllerBase):
    def __init__(self, n_time_steps, initial_time, final_time, time_unit="second"):
        super().__init__(initial_time, final_time, time_unit)
        self.n_time_steps = n_time_steps
        self.time_list = self.calculate_varying_times(self.fun_parabolic)
        self.dt = self.time_list[1] - self.time_list[0]
        self.step_counter = 0

    def fun_parabolic(self, t_array):
        return t_array ** 2

    def calculate_varying_times(self, fun):
        t_eq = np.linspace(self.t_initial, self.t_final, self.n_time_steps)
        y = fun(t_eq)
        f_min, f_max = np.min(t_eq), np.max(y)
        k = (t_eq.max() - t_eq.min()) / (f_max - f_min)
        return k * (y - f_min) + t_eq.min()

    def advance_time(self):
        self.step_counter += 1
        self.t = self.time_list[self.step_counter]
        self.dt = self.time_list[self.step_counter] - self.time_list[self.step_counter - 1]
