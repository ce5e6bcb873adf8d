"""Server / scheduler bootstrap (reference: python/mxnet/kvstore_server.py:41-86).

In the reference ``import mxnet`` in a process whose DMLC_ROLE is ``server`` or ``scheduler`` never returns to
user code: it runs the server loop and exits.  ``_init_kvstore_server_module()`` does the same here; call it
(or run ``python -m mxb200.kvstore_server``) in such a process.  tools/launch.py does this for you.
"""
import logging
import os
import sys


class KVStoreServer(object):
    def __init__(self, node):
        from .kvstore.dist import KVStoreDistServer
        self.server = KVStoreDistServer(node)
        self.init_logging = False

    def run(self):
        self.server.run()


def run_scheduler():
    from .kvstore.dist import PsNode, GROUP_ALL
    node = PsNode("scheduler")
    node.start()
    node.barrier(GROUP_ALL)      # start-up barrier
    node.barrier(GROUP_ALL)      # shutdown barrier (workers + servers + scheduler)
    node.stop()


def run_server():
    from .kvstore.dist import PsNode
    node = PsNode("server")
    node.start()
    KVStoreServer(node).run()


def _init_kvstore_server_module():
    role = os.environ.get("DMLC_ROLE", "worker")
    if role == "scheduler":
        logging.getLogger().info("Scheduler starting")
        run_scheduler()
        sys.exit(0)
    if role == "server":
        run_server()
        sys.exit(0)


if __name__ == "__main__":
    _init_kvstore_server_module()
