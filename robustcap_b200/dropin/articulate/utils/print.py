r"""Coloured print helpers evaluate.py star-imports (articulate/utils/print.py)."""


def _c(code):
    def f(*args, **kwargs):
        print('\033[%dm' % code + ' '.join(str(a) for a in args) + '\033[0m', **kwargs)
    return f


print_red, print_green, print_yellow, print_blue, print_purple, print_cyan = _c(31), _c(32), _c(33), _c(34), _c(35), _c(36)
