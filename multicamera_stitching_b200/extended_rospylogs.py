"""Minimal stand-in for the ``extended_rospylogs`` module the reference imports.

The reference's ``Stitcher``/``StitcherBase`` inherit from ``Debugger`` and log
through ``self.debugger(level, msg, log_type=...)``
(PostScripts/Stitcher/StitcherClass.py:16-17, :50, :125-127, :180).  That
module is not part of the reference repository (it lives in the robot's ROS
workspace), so this shim keeps the call sites and pickles working on top of the
standard :mod:`logging` package.
"""
import logging
import os

DEBUG_LEVEL_0 = 0
DEBUG_LEVEL_1 = 1
DEBUG_LEVEL_2 = 2
DEBUG_LEVEL_3 = 3
DEBUG_LEVEL_4 = 4

_LOG = logging.getLogger("multicamera_stitching_b200")

_LOG_FN = {
    "info": _LOG.info,
    "warn": _LOG.warning,
    "err": _LOG.error,
}


def _verbosity():
    try:
        return int(os.environ.get("MCS_DEBUG_LEVEL", DEBUG_LEVEL_0))
    except ValueError:
        return DEBUG_LEVEL_0


class Debugger(object):
    """Mix-in giving ``self.debugger(level, msg, log_type)``."""

    def debugger(self, level, msg, log_type="info"):
        if level > _verbosity():
            return
        _LOG_FN.get(log_type, _LOG.info)(msg)


def loginfo_cond(cond, msg):
    if cond:
        _LOG.info(msg)


def logerr_cond(cond, msg):
    if cond:
        _LOG.error(msg)


def update_debuggers(*_args, **_kwargs):
    return None
