"""ORACLE / TEST INFRASTRUCTURE ONLY.  ``greensim.logging.Filter`` (used by the reference's examples only:
``examples/network_simple/network_simple.py:9,26-34``): stamps log records with simulated time and process name."""
import logging as _logging

from . import Process


class Filter(_logging.Filter):

    def filter(self, record: _logging.LogRecord) -> bool:
        try:
            proc = Process.current()
            record.sim_time = proc._sim.now()
            record.sim_process = proc.local.name
        except TypeError:
            record.sim_time = -1.0
            record.sim_process = ""
        return True
