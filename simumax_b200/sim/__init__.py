"""Discrete-event pipeline simulator: coroutine engine, schedule generators, memory timeline, trace export."""
