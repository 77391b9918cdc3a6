"""Print this host's IPv4 address for the HOST file (role of the reference's experiment/ip.py;
uses the socket API because `ifaddr` is not installed here)."""
import socket

if __name__ == "__main__":
    s = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
    try:
        s.connect(("10.255.255.255", 1))
        print(s.getsockname()[0])
    except Exception:
        print("127.0.0.1")
    finally:
        s.close()
