"""Stand-in for the reference's logger.py (logger.py:4-33): same interface, but it
keeps lines in memory instead of appending under /root/reference/logs (read-only)."""


class Logger(object):
    def __init__(self, name=None, path=None):
        self.lines = []
        self.aux = []
        self.echo = False

    def write_text(self, txt, silent=False):
        self.lines.append(txt)
        if self.echo and not silent:
            print(txt)

    def write_text_aux(self, txt, silent=True):
        self.aux.append(txt)
