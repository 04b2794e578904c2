// TEST INFRASTRUCTURE ONLY -- a minimal, header-only stand-in for the slice of Qt5 that the
// reference's src/DVB_T2/*.cpp touches, so those translation units compile UNMODIFIED without Qt
// (the reference uses Qt only as QObject/signal/thread plumbing on this path).
// Threads, mutexes and wait conditions are no-ops: oracle/glue.cc turns every queued signal into a
// synchronous call, so the whole receiver runs on the caller's thread.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <iostream>

typedef unsigned int uint;

#define Q_OBJECT
#define signals public
#define slots
#define emit
#define Q_UNUSED(x) (void)x;
#define Q_DECLARE_METATYPE(T)
template <typename T> inline int qRegisterMetaType() { return 0; }
template <typename T> inline int qRegisterMetaType(const char*) { return 0; }

namespace Qt { enum ConnectionType { AutoConnection, DirectConnection, QueuedConnection }; }

class QString {
 public:
  std::string s;
  QString() {}
  QString(const char* c) : s(c) {}
  QString(const std::string& c) : s(c) {}
  static QString number(long long v) { return QString(std::to_string(v)); }
  static QString number(double v) { return QString(std::to_string(v)); }
  static QString number(int v) { return QString(std::to_string(v)); }
  static QString number(unsigned v) { return QString(std::to_string(v)); }
  static QString number(float v) { return QString(std::to_string(v)); }
  QString& operator+=(const QString& o) { s += o.s; return *this; }
  QString& operator+=(const char* o) { s += o; return *this; }
  bool operator==(const QString& o) const { return s == o.s; }
  bool operator!=(const QString& o) const { return s != o.s; }
  std::string toStdString() const { return s; }
};
inline QString operator+(const QString& a, const QString& b) { return QString(a.s + b.s); }
inline QString operator+(const QString& a, const char* b) { return QString(a.s + b); }
inline QString operator+(const char* a, const QString& b) { return QString(std::string(a) + b.s); }

struct QDebugSink {
  template <typename T> QDebugSink& operator<<(const T&) { return *this; }
  QDebugSink& operator<<(const QString&) { return *this; }
};
inline QDebugSink qDebug() { return QDebugSink(); }

class QThread;
class QObject {
 public:
  explicit QObject(QObject* = nullptr) {}
  virtual ~QObject() {}
  template <typename... A> static bool connect(A&&...) { return true; }
  template <typename... A> static bool disconnect(A&&...) { return true; }
  void moveToThread(QThread*) {}
  void deleteLater() {}
};

class QMutex {
 public:
  void lock() {}
  void unlock() {}
  bool try_lock() { return true; }
  bool tryLock(int = 0) { return true; }
};
class QWaitCondition {
 public:
  void wakeOne() {}
  void wakeAll() {}
  bool wait(QMutex*, unsigned long = 0) { return true; }
};
class QThread : public QObject {
 public:
  enum Priority { IdlePriority, LowestPriority, LowPriority, NormalPriority, HighPriority, HighestPriority,
                  TimeCriticalPriority, InheritPriority };
  explicit QThread(QObject* = nullptr) {}
  void start(Priority = InheritPriority) {}
  bool isRunning() const { return false; }
  bool wait(unsigned long = 0) { return true; }
  void quit() {}
  void finished() {}
  static void msleep(unsigned long) {}
};

// ---- the TS sink: everything bb_de_header writes (UDP datagram or file) lands here ----
struct OracleTsSink {
  std::vector<char> bytes;
  std::vector<int> datagram_len;
  static OracleTsSink& get() { static OracleTsSink s; return s; }
};

struct QHostAddress { enum SpecialAddress { LocalHost }; };
class QIODevice { public: enum OpenModeFlag { ReadOnly = 1, WriteOnly = 2 }; };
class QFile : public QIODevice {
 public:
  QString name;
  bool opened = false;
  explicit QFile(const QString& n) : name(n) {}
  bool open(int) { opened = true; return true; }
  bool isOpen() const { return opened; }
  void close() { opened = false; }
  QString errorString() const { return QString("error"); }
};
class QDataStream {
 public:
  enum Version { Qt_5_15 = 19 };
  void setVersion(int) {}
  void setDevice(QIODevice*) {}
  int writeRawData(const char* p, int n) {
    auto& s = OracleTsSink::get(); s.bytes.insert(s.bytes.end(), p, p + n); s.datagram_len.push_back(n); return n;
  }
};
class QUdpSocket : public QObject {
 public:
  explicit QUdpSocket(QObject* = nullptr) {}
  long long writeDatagram(const char* p, long long n, QHostAddress::SpecialAddress, unsigned short) {
    auto& s = OracleTsSink::get(); s.bytes.insert(s.bytes.end(), p, p + n); s.datagram_len.push_back((int)n); return n;
  }
  bool isOpen() const { return false; }
  void close() {}
};
struct QMessageBox { template <typename... A> static int information(A&&...) { return 0; } };
class QTextStream {};
class QCoreApplication { public: static void processEvents() {} };
class QApplication : public QCoreApplication {};
